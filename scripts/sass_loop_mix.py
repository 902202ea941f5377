"""Instruction mix of a kernel's hottest loop (largest backward branch) from cuobjdump -sass.
usage: python scripts/sass_loop_mix.py <lib.so> <kernel-name-substring> [entries_per_iteration]"""
import re, subprocess, sys, collections
lib, name = sys.argv[1], sys.argv[2]
per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs:
    if name not in f.split("\n")[0]:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r"BRA\s+(?:U\w+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    print(f.split("\n")[0])
    if not best:
        print("no loop"); continue
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    mix = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        mix[t.split()[0].split(".")[0]] += 1
    print(f"loop 0x{best[0]:x}..0x{best[1]:x}: {len(body)} instructions, {len(body)/per:.1f} per entry")
    print("  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common()))
