"""Small end-to-end run for compute-sanitizer (memcheck): single-GPU step path, density field,
estimate_densities, slab phases with world 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api, slab
ref = np.load(os.path.join(H.GOLDEN, "ref_jitter_two_blocks.npz"))
g = api.Solver(api.default_params(rest_density=700.0))
g.upload(ref["pos"], ref["vel"]); g.estimate_densities(); g.step(2)
g.density_at(ref["pos"][:100]); g.step(1); g.neighbor_digest(); g.neighbors(); g.download()
os.environ["PBF_GRAPH"] = "0"
g = api.Solver(api.default_params(rest_density=700.0)); g.upload(ref["pos"], ref["vel"]); g.step(2); g.download()
s = slab.SlabSolver(api.default_params(rest_density=700.0), 0, 1)
s.upload_local(ref["pos"], ref["vel"]); s.step(2); s.download_local(); s.neighbor_digest()
print("sanitize run ok")
