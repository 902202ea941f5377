"""Dev script: pbf_debug_probe on the bench state (C4 dam break, or --block nx ny nz): what each ingredient of the gather kernels costs.
    python scripts/probe_gathers.py [--block 400 200 200] [--steps 5] [--settle 0] [--variants 0 1 2 3 4 5 6 7 8 9]"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from fluid_b200 import api
import bench

NAMES = {0: "lists only (no gather)", 1: "gather 4 B", 2: "gather 8 B", 3: "gather 16 B", 4: "gather 8 B + 4 B", 5: "gather 32 B",
         6: "k_lambda (shipped)", 7: "lambda arithmetic, no gather", 8: "lambda, 8-byte records, scalar", 9: "lambda, 8-byte records, packed fp32x2",
         10: "lambda, 8-byte records, packed, tuned", 11: "as 10, 6 CTAs/SM (40 registers)", 12: "as 10, next row's records in flight",
         13: "delta-p sum, 8-byte pos + 4-byte lambda, packed, pipelined", 14: "as 12, 6 CTAs/SM", 15: "as 13, 6 CTAs/SM",
         16: "k_lambda arithmetic, float2 (x, y) + float z gathers", 18: "delta-p sum, float2 (x, y) + float2 (z, lambda) gathers", 19: "delta-p sum, one float4 gather (as shipped)", 20: "k_lambda arithmetic, two particles per lane (slices 2w, 2w+1 interleaved)",
         22: "k_lambda arithmetic, two lanes per particle (lanes 2p, 2p+1 take alternate list rows)", 23: "k_lambda arithmetic, two lanes per particle (lanes p, p+16)",
         24: "k_lambda arithmetic, 16-bit delta list entries (8 per 16-byte row, escapes from a side array)"}

ap = argparse.ArgumentParser()
ap.add_argument("--block", type=int, nargs=3, default=[400, 200, 200])
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variants", type=int, nargs="*", default=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
ap.add_argument("--out", default=None)
args = ap.parse_args()
nx, ny, nz = args.block
box_max = (max(120.0, 0.3 * nx), 30.0, 0.1 * nz + 0.1)
prm = api.default_params(rest_density=700.0, iterations=12, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
g = api.Solver(prm)
pos, vel = bench.block_f32(nx, ny, nz)
g.upload(pos, vel); g.step(args.steps)
n = len(pos)
lib = g.lib
lib.pbf_debug_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]
res = {"particles": n, "steps": args.steps, "variants": {}}
ref = None
for v in args.variants:
    ms = C.c_double()
    out = np.zeros((n, 4), dtype=np.float32) if v in (6, 8, 9, 10, 12, 16, 20, 22, 23, 24) or v // 10 in (10, 12) else None
    rc = lib.pbf_debug_probe(g.h, v, args.reps, C.byref(ms), out.ctypes.data_as(C.c_void_p) if out is not None else None)
    if rc != 0:
        print(v, "failed", rc, lib.pbf_last_error(g.h).decode()); continue
    e = {"name": NAMES.get(v, NAMES.get(v // 10, str(v)) + f", {v % 10} CTAs/SM" if v >= 100 else str(v)), "ms": ms.value}
    if v == 6:
        ref = out[:, 3].astype(np.float64).copy()          # lambda in .w
    elif out is not None and ref is not None:
        lam = out[:, 1].astype(np.float64)
        e["lambda_max_abs_diff_rel_to_max"] = float(np.abs(lam - ref).max() / np.abs(ref).max())
        e["lambda_p99_abs_diff_rel_to_max"] = float(np.percentile(np.abs(lam - ref), 99) / np.abs(ref).max())
        if os.environ.get("PROBE_WORST"):
            d = np.abs(lam - ref); w = np.argsort(d)[-5:]
            print("   worst:", [(int(i), float(ref[i]), float(lam[i]), float(out[i, 0])) for i in w], "n bad (>1e-4 max):", int((d > 1e-4 * np.abs(ref).max()).sum()))
    res["variants"][v] = e
    print(f"variant {v:2d}  {e['ms']:7.3f} ms  {e['name']}  " + " ".join(f"{k}={val:.3e}" for k, val in e.items() if k.startswith("lambda")), flush=True)
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
