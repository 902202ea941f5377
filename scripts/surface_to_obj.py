#!/usr/bin/env python
"""Convert a surface file written by `pbf_run --surface f` (or `oracle/_ref/ref_harness --surface f`: int64 count, then 18
doubles per triangle = p1 p2 p3 n1 n2 n3) into a Wavefront .obj with per-vertex normals, for any mesh viewer.
usage: surface_to_obj.py surf.bin out.obj"""
import sys
import numpy as np


def main(src, dst):
    raw = np.fromfile(src, dtype=np.uint8)
    nt = int(raw[:8].view(np.int64)[0])
    t = raw[8:8 + 144 * nt].view(np.float64).reshape(nt, 18)
    with open(dst, "w") as f:
        f.write(f"# {nt} marching-cubes triangles (Particles::getSurfacePrims)\n")
        for p in t[:, :9].reshape(-1, 3):
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for n in t[:, 9:].reshape(-1, 3):
            f.write("vn %.6g %.6g %.6g\n" % tuple(n))
        for k in range(nt):
            a = 3 * k + 1
            f.write(f"f {a}//{a} {a + 1}//{a + 1} {a + 2}//{a + 2}\n")
    print(f"{dst}: {nt} triangles")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    main(sys.argv[1], sys.argv[2])
