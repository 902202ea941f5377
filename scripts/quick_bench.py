"""Quick per-kernel timing of the step on a pgen-style dam-break block (dev tool, not bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluid_b200 import api

def block(nx, ny, nz, jitter=0.001):
    i, j, k = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(nz, dtype=np.float32), indexing="ij")
    pos = np.stack([0.1 + 0.1 * i, 0.1 + 0.1 * j, 0.1 + 0.1 * k], axis=-1).reshape(-1, 3).astype(np.float64)
    if jitter:
        pos += np.random.default_rng(1234).uniform(-jitter, jitter, size=pos.shape)
    vel = np.tile(np.array([0.0, -1.0, 0.0]), (pos.shape[0], 1))
    return pos, vel

def main():
    nx, ny, nz = (int(a) for a in sys.argv[1:4])
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 12
    box_max = (max(30.0, 0.3 * nx), max(15.0, 0.15 * ny), 0.1 * nz + 0.1)
    if os.environ.get("QB_CONFINED"):      # the wide-tank geometry of the N > 1 bench: walls right behind both x-faces of the block
        box_max = (0.1 * nx + 0.1, 30.0, 0.1 * nz + 0.1)
    p = api.default_params(rest_density=700.0, iterations=iters, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
    s = api.Solver(p)
    if os.environ.get("QB_OBSTACLES"):      # two spheres and a cuboid standing in the block (particles inside them are left out)
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import helpers as H
        sph = np.array([[0.05 * nx, 0.0, 0.05 * nz, 0.04 * nz], [0.08 * nx, 0.03 * ny, 0.03 * nz, 0.02 * nz]])
        lo, hi = (0.02 * nx, 0.0, 0.06 * nz), (0.03 * nx, 0.05 * ny, 0.09 * nz)
        s.set_obstacle_spheres(sph); s.set_obstacle_triangles(H.box_mesh(lo, hi))
    mesh_c = None
    if os.environ.get("QB_MESH"):           # a tessellated sphere (2 * nlon * (nlat - 1) triangles, nlon = 2 nlat) standing in the block: device BVH
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import helpers as H
        nlat = int(os.environ["QB_MESH"])
        mesh_c = (0.05 * nx, 0.0, 0.05 * nz, 0.04 * nz)
        t0 = time.time(); tris = H.uv_sphere_mesh(mesh_c[:3], mesh_c[3], nlat, 2 * nlat); t_gen = time.time() - t0
        t0 = time.time(); s.set_obstacle_triangles(tris); t_set = time.time() - t0
        print(f"mesh: {len(tris)} triangles, generated in {t_gen:.1f} s, pbf_set_obstacle_triangles (host BVH build + upload) {t_set:.3f} s", file=sys.stderr)
    pos, vel = block(nx, ny, nz)
    if mesh_c is not None:
        keep = np.linalg.norm(pos - np.array(mesh_c[:3]), axis=1) > mesh_c[3] + 0.02
        pos, vel = pos[keep], vel[keep]
    if os.environ.get("QB_OBSTACLES"):
        keep = np.ones(len(pos), dtype=bool)
        for c in sph:
            keep &= np.linalg.norm(pos - c[:3], axis=1) > c[3] + 0.02
        keep &= ~np.all((pos > np.array(lo) - 0.02) & (pos < np.array(hi) + 0.02), axis=1)
        pos, vel = pos[keep], vel[keep]
    n = pos.shape[0]
    t0 = time.time(); s.upload(pos, vel); t_up = time.time() - t0
    s.step(2)   # warm-up
    s.profile_enable(True)
    s.step(steps)
    a, b, ms = s.stats()
    prof = s.profile()
    s.profile_enable(False)
    s.step(steps)
    a, b, ms2 = s.stats()
    d, c = s.neighbor_digest()
    t0 = time.time(); P, V, R = s.download(); t_down = time.time() - t0
    out = dict(n=n, iters=iters, steps=steps, ms_per_step_profiled=ms / steps, ms_per_step=ms2 / steps,
               particle_iter_per_s=n * iters * steps / (ms2 * 1e-3), avg_rho=(a, b), mean_nbrs=float(c.mean()), max_nbrs=int(c.max()),
               upload_s=t_up, download_s=t_down,
               kernels={k: dict(ms_per_step=v[0] / steps, launches_per_step=v[1] / steps) for k, v in prof.items() if v[1]})
    if mesh_c is not None:
        out["inside_mesh"] = int((np.linalg.norm(P - np.array(mesh_c[:3]), axis=1) < mesh_c[3] * 0.999).sum())
    print(json.dumps(out, indent=1))

main()
