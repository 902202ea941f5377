"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck, scripts/sanitize.sh): single-GPU step path
(plain launches and graph replay), streaming read-back, density field, estimate_densities, slab phases with world 1,
obstacles and the surfacer."""
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
# streaming read-back on a second stream racing the next step (pbf_set_readback), plain launches and graph replay
for graph in ("0", "1"):
    os.environ["PBF_GRAPH"] = graph
    g = api.Solver(api.default_params(rest_density=700.0)); g.upload(ref["pos"], ref["vel"])
    P = np.empty((g.n, 3)); V = np.empty((g.n, 3)); R = np.empty(g.n)
    g.pin(P, V, R); g.set_readback(P, V, R)
    for _ in range(3):
        g.step(1, sync=False); g.step(1, sync=False); g.sync()
    g.density_at(ref["pos"][:50]); g.step(1, sync=False); g.estimate_densities(); g.step(2)
    P2, V2, R2 = g.download()
    assert np.array_equal(P, P2) and np.array_equal(V, V2) and np.array_equal(R, R2)
os.environ["PBF_GRAPH"] = "0"
s = slab.SlabSolver(api.default_params(rest_density=700.0), 0, 1)
s.upload_local(ref["pos"], ref["vel"]); s.step(2); s.download_local(); s.neighbor_digest()
# obstacle spheres + a triangle mesh through the device BVH, and the marching-cubes surfacer
g = api.Solver(api.default_params(rest_density=700.0))
keep = ref["pos"][:, 1] >= 0.75
mesh = np.concatenate([H.uv_sphere_mesh((-0.5, 0.32, 0.5), 0.3, 16, 32), H.heightfield_mesh(-1.05, 1.05, -1.05, 1.05, 24, 24)])
g.set_obstacle_triangles(mesh); g.set_obstacle_spheres(np.array([[0.5, 0.4, -0.5, 0.2]]))
g.upload(ref["pos"][keep], ref["vel"][keep]); g.step(25)
t = g.extract_surface(700.0); g.step(1); t2 = g.extract_surface(700.0, step=0.07)
print("surface triangles", len(t), len(t2))
print("sanitize run ok")
