#!/usr/bin/env python
"""Particle files for the `-p` option (the reference's particles/pgen.py writes the two-block spheres_p.xml scene through an
in-memory ElementTree; this one streams, so multi-million-particle dam-break blocks are fine): an axis-aligned lattice
block, index order x outer / y / z inner, spacing 0.1, first particle 0.1 from the walls, v = (0, -1, 0), rho0 = 700 —
the synthetic recipe of BASELINE configs C3-C5.  `--two-blocks` writes exactly the reference's pgen.py scene instead.

usage: pgen_block.py nx ny nz out.xml [--jitter 0.001] [--seed 1234] [--density 700] [--two-blocks]"""
import argparse
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", type=int, nargs=3)
    ap.add_argument("out")
    ap.add_argument("--jitter", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--density", type=float, default=700.0)
    ap.add_argument("--two-blocks", action="store_true")
    a = ap.parse_args()
    with open(a.out, "w") as f:
        f.write("<?xml version=\"1.0\"?>\n<particles>\n  <density>%r</density>\n  <ps>\n" % a.density)
        if a.two_blocks:                                  # particles/pgen.py:54-61
            rows = [(0.1 * i - 1, 0.1 * j, 1 - 0.1 * k) for i in range(1, 10) for j in range(1, 14) for k in range(1, 10)]
            rows += [(1 - 0.1 * i, 0.1 * j, 0.1 * k - 1) for i in range(1, 10) for j in range(1, 14) for k in range(1, 10)]
            chunks = [np.array(rows)]
        else:
            nx, ny, nz = a.dims
            rng = np.random.default_rng(a.seed)
            j, k = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")

            def gen():
                for i in range(nx):                       # one x-layer at a time; jitter drawn in particle-index order x, y, z
                    p = np.stack([np.full(j.shape, 0.1 + 0.1 * i), 0.1 + 0.1 * j, 0.1 + 0.1 * k], -1).reshape(-1, 3)
                    if a.jitter > 0:
                        p = p + rng.uniform(-a.jitter, a.jitter, size=p.shape)
                    yield p
            chunks = gen()
        n = 0
        for p in chunks:
            f.write("".join("    <particle>\n      <pos>%r %r %r</pos>\n      <v>0 -1 0</v>\n    </particle>\n" % (float(x), float(y), float(z)) for x, y, z in p))
            n += len(p)
        f.write("  </ps>\n</particles>\n")
    print(f"{a.out}: {n} particles")


if __name__ == "__main__":
    main()
