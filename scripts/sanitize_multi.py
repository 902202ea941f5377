"""Small multi-slab run for compute-sanitizer (scripts/sanitize.sh multi): three slabs behind ONE pbf_multi handle sharing the
GPU (peer mode: peer stores into the neighbours' buffers, flag hand-overs, device-side ranges), a block flowing along a long
tank so that particles migrate and the slabs are re-balanced, obstacle sphere on a slab boundary; result == one handle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api
nx, ny, nz = 48, 8, 6
pos, vel = H.lattice_block(nx, ny, nz, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(2.5, -1.0, 0.0), jitter=0.001, seed=5)
box_max = (0.3 * nx + 0.3, 3.0, 0.1 * nz + 0.3)
prm = dict(rest_density=700.0, iterations=4, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
sph = np.array([[1.65, 0.3, 0.4, 0.25]])
m = api.MultiSolver(api.default_params(**prm), devices=[0, 0, 0])
m.set_rebalance(2, 1.02)
m.set_obstacle_spheres(sph)
m.upload(pos, vel); m.step(12)
P, V, R = m.download()
plan = m.plan()
g = api.Solver(api.default_params(**prm)); g.set_obstacle_spheres(sph); g.upload(pos, vel); g.step(12)
P1, V1, R1 = g.download()
assert np.array_equal(P, P1) and np.array_equal(V, V1) and np.array_equal(R, R1), "slabs differ from one handle"
print("re-balancings", plan[2] if isinstance(plan, tuple) else plan)
print("sanitize multi ok")
