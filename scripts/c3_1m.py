"""BASELINE config C3: dam-break 1M particles (100^3 lattice, box (30, 15, 10.1)) on one B200 next to the CPU oracle port
(fp64, uniform grid, OpenMP, all host cores, 3 steps).  Report tool; bench.py stays on C4."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api

pos, vel = H.lattice_block(100, 100, 100, jitter=0.001)
box = dict(box_min=(0, 0, 0), box_max=(30.0, 15.0, 10.1), y_light=15.0, z_front=10.1)
out = {"config": "C3 dam-break 100x100x100 = 1000000 particles, spacing 0.1, H 0.3, rho0 700, box (30, 15, 10.1)", "gpu": {}, "cpu": {}}
for iters in (12, 4):
    g = api.Solver(api.default_params(rest_density=700.0, iterations=iters, **box))
    g.upload(pos, vel); g.step(3)
    g.step(20); ms = g.stats()[2] / 20
    out["gpu"][f"I={iters}"] = {"ms_per_step": ms, "updates_per_s": len(pos) * iters / (ms * 1e-3), "launches_per_step": 12 + 2 * iters}
o = H.Oracle(H.default_params(rest_density=700.0, iterations=12, xsph_mode=H.XSPH_JACOBI, **box), 64, H.COLLIDE_BOX, H.SEARCH_GRID)
o.upload(pos, vel)
t0 = time.perf_counter(); o.step(3); dt = (time.perf_counter() - t0) / 3
out["cpu"] = {"kind": "oracle port (fp64, grid, OpenMP)", "threads": H.oracle_lib().oracle_max_threads(), "iterations": 12, "steps": 3,
              "ms_per_step": dt * 1e3, "updates_per_s": len(pos) * 12 / dt}
out["note"] = "the unmodified reference cannot run this config: its neighbour search is O(N^2) (0.55 s/step at 8k particles => hours per step at 1M)"
print(json.dumps(out, indent=1))
