"""Dev check: neighbour sets of EVOLVED (disordered) states vs the fp32 oracle, and the spread of the 25-step kinetic
energy over in-cell orderings (PBF_ZSUB) — how sensitive that aggregate is to fp32 summation order."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api
ref = np.load(os.path.join(H.GOLDEN, "ref_jitter_two_blocks.npz"))
pos, vel, rho0 = ref["pos"], ref["vel"], float(ref["rho0"])
out = {}
g = api.Solver(api.default_params(rest_density=rho0)); g.upload(pos, vel)
bad = 0
for k in range(1, 41):
    g.step(1)
    if k % 5 == 0:
        P, V, R = g.download()
        g0 = api.Solver(api.default_params(rest_density=rho0, iterations=0)); g0.upload(P, V); g0.step(1)
        o = H.Oracle(H.default_params(rest_density=rho0, iterations=0, xsph_mode=H.XSPH_JACOBI), 32, H.COLLIDE_BOX, H.SEARCH_GRID)
        o.upload(P, V); o.step(1)
        dg, cg = g0.neighbor_digest(); do, co = o.digest()
        nb = int((dg != do).sum()); bad += nb
        out[f"step{k}"] = {"digest_mismatch": nb, "mean_nbrs": float(cg.mean()), "ke": float(0.5 * (V ** 2).sum())}
out["digest_mismatch_total"] = bad
o64 = H.Oracle(H.default_params(rest_density=rho0, xsph_mode=H.XSPH_JACOBI), 64, H.COLLIDE_BOX, H.SEARCH_GRID); o64.upload(pos, vel)
o32 = H.Oracle(H.default_params(rest_density=rho0, xsph_mode=H.XSPH_JACOBI), 32, H.COLLIDE_BOX, H.SEARCH_GRID); o32.upload(pos, vel)
ke64, ke32 = [], []
for k in range(30):
    o64.step(1); o32.step(1)
    ke64.append(float(0.5 * (o64.download()[1] ** 2).sum())); ke32.append(float(0.5 * (o32.download()[1] ** 2).sum()))
out["ke_oracle64"] = ke64[19:30]; out["ke_oracle32"] = ke32[19:30]
print(json.dumps(out))
