"""Dev script: the flowing dam break of tests/test_gpu_multi.py with PBF_MULTI_TRACE=1 (prints every re-balancing decision)."""
import os, sys
os.environ.setdefault("PBF_MULTI_TRACE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import torch
from fluid_b200 import api
from helpers import lattice_block

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = torch.cuda.device_count()
devs = [d % n for d in range(world)] if n >= world else [0] * world
if os.environ.get("PBF_DEVS"):
    devs = [int(d) for d in os.environ["PBF_DEVS"].split(",")]
nx, ny, nz = 96, 40, 8
pos, vel = lattice_block(nx, ny, nz, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(0.0, 0.0, 0.0), jitter=0.001, seed=5)
box_max = (30.3, 6.0, 0.1 * nz + 0.3)
prm = dict(rest_density=700.0, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
m = api.MultiSolver(api.default_params(**prm), devices=devs)
m.set_rebalance(2, 1.05)
m.upload(pos, vel)
for k in range(10):
    m.step(30)
    b, owned, nreb = m.plan()
    P, _, _ = m.download()
    print(k, b.tolist(), owned.tolist(), nreb, "x range", P[:, 0].min(), P[:, 0].max(), flush=True)
