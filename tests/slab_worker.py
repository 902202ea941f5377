"""Worker for the slab-decomposition tests: run under torchrun (or directly for world 1).

Runs K steps of a wide jittered block in slab mode on WORLD ranks, gathers the particles by global
id and (rank 0) compares them with a single-handle run of the same input through pbf_step: the two
must agree BIT FOR BIT (same global cell grid, same in-cell order, same summation order).
Prints one JSON line on rank 0.   --backend gloo lets several ranks share one GPU (host-staged halos).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def scene(nx, ny, nz, seed=11):
    from helpers import lattice_block
    pos, vel = lattice_block(nx, ny, nz, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(0.0, -1.0, 0.0), jitter=0.001, seed=seed)
    rng = np.random.default_rng(seed + 1)
    vel = vel + rng.normal(0.0, 0.3, size=vel.shape) + np.array([1.5, 0.0, 0.0]) * np.sin(pos[:, :1] * 2.0)   # x-motion => migration
    return pos, vel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--dims", type=int, nargs=3, default=[96, 20, 20])
    ap.add_argument("--iterations", type=int, default=12)
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--spheres", action="store_true", help="obstacle spheres on the floor, one of them across a slab boundary")
    ap.add_argument("--mesh", action="store_true", help="a tessellated sphere (device BVH) across the slab boundary")
    ap.add_argument("--transport", default=None, help="p2p (CUDA IPC, the library's peer mode) | nccl | staged; default: SlabSolver's choice")
    ap.add_argument("--flow", action="store_true", help="shallow block pushed along x in a long tank: the slabs have to be re-balanced")
    ap.add_argument("--rebalance-every", type=int, default=8)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from fluid_b200 import api, slab

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    nx, ny, nz = args.dims
    box_max = (0.1 * nx + 0.4, 0.1 * ny + 2.0, 0.1 * nz + 0.3)
    prm = dict(rest_density=700.0, iterations=args.iterations, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
    pos, vel = scene(nx, ny, nz)
    if args.flow:
        from helpers import lattice_block
        pos, vel = lattice_block(nx, ny, nz, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(3.0, -1.0, 0.0), jitter=0.001, seed=5)
        box_max = (0.3 * nx + 0.3, 4.0, 0.1 * nz + 0.3)
        prm.update(box_max=box_max, y_light=box_max[1], z_front=box_max[2])
    # obstacle spheres are global scene data: every rank sets the same list.  The first one sits on the slab
    # boundary of a 2-rank run; particles that would start inside a sphere are left out of the block.
    spheres = np.array([[0.05 * nx, 0.6, 0.05 * nz, 0.7], [0.025 * nx, 1.2, 0.03 * nz, 0.5]]) if args.spheres else np.zeros((0, 4))
    for c in spheres:
        keep = np.linalg.norm(pos - c[:3], axis=1) > c[3] + 0.02
        pos, vel = pos[keep], vel[keep]
    mesh = None
    if args.mesh:
        import helpers as H
        mc = np.array([0.05 * nx + 0.13, 0.5, 0.05 * nz]); mr = 0.6
        mesh = H.uv_sphere_mesh(mc, mr, 20, 40)
        keep = np.linalg.norm(pos - mc, axis=1) > mr + 0.02
        pos, vel = pos[keep], vel[keep]
    n = pos.shape[0]
    # rank r starts with the r-th contiguous chunk of the x-outer lattice order (roughly its slab)
    per = n // world
    lo = rank * per; hi = n if rank == world - 1 else (rank + 1) * per
    s = slab.SlabSolver(api.default_params(**prm), rank, world, device=local, transport=args.transport, rebalance_every=args.rebalance_every)
    s.set_obstacle_spheres(spheres)
    if mesh is not None:
        s.set_obstacle_triangles(mesh)
    s.upload_local(pos[lo:hi], vel[lo:hi], id_offset=lo)
    s.step(args.steps); s.sync()
    P, V, R, I, d, c = s.gather_all()
    a_first, a_final = s.stats()
    owned = [0] * world
    if world > 1:
        dist.all_gather_object(owned, int(s.n_owned()))
    else:
        owned = [int(s.n_owned())]
    out = {"world": world, "n": int(n), "steps": args.steps, "bounds": list(s.bounds), "col_bounds": list(map(int, s.col_bounds)),
           "transport": s.transport, "p2p_fallback_reason": getattr(s, "p2p_fallback_reason", None), "n_rebalances": s.n_rebalances, "owned": owned}
    if rank == 0:
        g = api.Solver(api.default_params(**prm), device=local)
        g.set_obstacle_spheres(spheres)
        if mesh is not None:
            g.set_obstacle_triangles(mesh)
        g.upload(pos, vel); g.step(args.steps)
        Pg, Vg, Rg = g.download()
        dg, cg = g.neighbor_digest()
        a, b, _ = g.stats()
        out.update({
            "ids_ok": bool(np.array_equal(I, np.arange(n, dtype=np.uint32))),
            "pos_equal": bool(np.array_equal(P, Pg)), "vel_equal": bool(np.array_equal(V, Vg)), "rho_equal": bool(np.array_equal(R, Rg)),
            "digest_equal": bool(np.array_equal(d, dg) and np.array_equal(c, cg)),
            "max_dpos": float(np.abs(P - Pg).max()) if len(P) == len(Pg) else None,
            "avg_rho_slab": [a_first, a_final], "avg_rho_single": [a, b],
            "finite": bool(np.isfinite(P).all()),
            "inside_mesh": int((np.linalg.norm(P - mc, axis=1) < mr * 0.99).sum()) if mesh is not None else None,
            "min_sphere_gap": float(min((np.linalg.norm(P - c[:3], axis=1).min() - c[3]) for c in spheres)) if len(spheres) else None,
        })
        print("SLAB_RESULT " + json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
