"""Dev check: predicted positions and neighbour sets of a 110k-particle dam break after 15..60 free GPU steps (disordered states)
against the fp32 oracle, bit for bit.  Result of the final build: profiles/r01_evolved_big_check.json."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api
pos, vel = H.lattice_block(48, 48, 48, jitter=0.001)
box = dict(box_min=(0, 0, 0), box_max=(12.0, 8.0, 4.9), y_light=8.0, z_front=4.9)
g = api.Solver(api.default_params(rest_density=700.0, **box)); g.upload(pos, vel)
out = {}
for k in range(1, 61):
    g.step(1)
    if k % 15 == 0:
        P, V, R = g.download()
        g0 = api.Solver(api.default_params(rest_density=700.0, iterations=0, **box)); g0.capture(True); g0.upload(P, V); g0.step(1)
        o = H.Oracle(H.default_params(rest_density=700.0, iterations=0, xsph_mode=H.XSPH_JACOBI, **box), 32, H.COLLIDE_BOX, H.SEARCH_GRID)
        o.upload(P, V); o.step(1)
        dg, cg = g0.neighbor_digest(); do, co = o.digest()
        out[f"step{k}"] = {"digest_mismatch": int((dg != do).sum()), "count_mismatch": int((cg != co).sum()), "mean_nbrs": float(cg.mean()), "max_nbrs": int(cg.max()),
                           "xpred_equal": bool(np.array_equal(g0.array(H.ARRAY_XPRED), o.array(H.ARRAY_XPRED)))}
print(json.dumps(out, indent=1))
