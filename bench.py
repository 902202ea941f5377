#!/usr/bin/env python
"""bench.py — particle-iteration updates/sec of the PBF step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one Particles::timeStep (predict, neighbour search, I solver iterations, XSPH +
vorticity, commit) over the whole synthetic particle set.  metric = N * I * steps / time.

Workloads (BASELINE.json configs):
  N=1   C4: dam-break block 400x200x200 = 16M particles, spacing 0.1, H 0.3, rho0 700, v0 (0,-1,0),
        box (0,0,0)->(120,30,20.1), 12 iterations, vorticity + XSPH on.
  N>1   C5-style wide tank, weak scaling: 400*N x 200 x 200 particles (16M per GPU; N=8 is the
        128M config), box (0,0,0)->(40*N+0.1,30,20.1), x-slab per rank, ghost halo exchange and
        migration over NCCL (fluid_b200/slab.py).
Inputs are larger than L2 (126 MB): 16M particles = 256 MB per float4 array, so no L2 flush is
needed between timed iterations.

Keys beyond the base contract: "roofline" (dominant kernel, algorithmic bytes / CUDA-event time /
measured HBM peak), "cpu_baseline" (oracle port on the host cores, bounded sample), "e2e" (host
buffers in and out every step through the C ABI), "gpu_launches", "clocks", "kernels".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

ITERATIONS = 12
BLOCK_PER_GPU = (400, 200, 200)          # 16M particles
SPACING, RHO0 = 0.1, 700.0
# algorithmic HBM bytes per particle: per iteration lambda 20 + delta 36; fixed 328
BYTES_LAMBDA, BYTES_DELTA, BYTES_FIXED = 20, 36, 328


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def block_f32(nx, ny, nz, x0=0.1, jitter=0.001, seed=1234, chunk=0, nchunks=1):
    """pgen-style lattice block (particles/pgen.py:54-61 generalised): index order x outer, y, z inner;
    jitter U(-j, j) from default_rng(seed).  Returns float64 [n,3] pos and vel."""
    xs = np.arange(nx, dtype=np.float64)
    if nchunks > 1:
        per = nx // nchunks
        xs = xs[chunk * per:(chunk + 1) * per]
    i, j, k = np.meshgrid(xs, np.arange(ny, dtype=np.float64), np.arange(nz, dtype=np.float64), indexing="ij")
    pos = np.stack([x0 + SPACING * i, 0.1 + SPACING * j, 0.1 + SPACING * k], axis=-1).reshape(-1, 3)
    if jitter:
        rng = np.random.default_rng(seed + chunk)
        pos += rng.uniform(-jitter, jitter, size=pos.shape)
    vel = np.zeros_like(pos); vel[:, 1] = -1.0
    return pos, vel


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---- CPU legs (the ONLY place bench.py touches oracle/) ---------------------------------------------

def cpu_baseline_port(target_seconds=12.0, min_side=32):
    """The oracle (CPU restatement, uniform grid, OpenMP, fp64, Jacobi XSPH) on all host cores, on a
    bounded sub-block of the same dam-break lattice, sized for ~10-30 s of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    cores = H.oracle_lib().oracle_max_threads()
    dims = (32, 32, 32)
    kw = dict(rest_density=RHO0, iterations=ITERATIONS, box_min=(0, 0, 0), box_max=(120.0, 30.0, 20.1), y_light=30.0, z_front=20.1,
              xsph_mode=H.XSPH_JACOBI)

    def run(dims, steps):
        pos, vel = block_f32(*dims)
        o = H.Oracle(H.default_params(**kw), 64, H.COLLIDE_BOX, H.SEARCH_GRID)
        o.upload(pos, vel)
        t0 = time.perf_counter(); o.step(steps); dt = time.perf_counter() - t0
        return pos.shape[0] * ITERATIONS * steps / dt, dt
    rate, dt = run(dims, 1)                                   # calibration
    n_target = rate * target_seconds / (ITERATIONS * 2)      # 2 steps
    side = int(max(min_side, min(160, round(n_target ** (1 / 3) / 8) * 8)))
    dims = (side, side, side)
    rate, dt = run(dims, 2)
    return {"value": rate, "unit": "particle-iteration updates/s", "cores": cores, "kind": "port",
            "sample": f"{dims[0]}x{dims[1]}x{dims[2]} sub-block ({dims[0] * dims[1] * dims[2]} particles) of the dam-break lattice, 2 steps, "
                      f"{ITERATIONS} iterations, fp64 oracle (uniform grid + OpenMP), {dt:.1f} s"}


def reference_arm(args):
    """--impl reference: the UNMODIFIED reference solver (oracle/_ref/ref_harness, built from
    /root/reference by oracle/Makefile) on the host.  It is single-threaded and its neighbour search
    is O(N^2) (particles.cpp:258-265), so a step is a bounded sample: a 20x14x20 block (5600 particles)
    of the same lattice inside the reference's hard-coded Cornell box.  Falls back to the oracle
    port when the compiled reference is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    binp = H.ref_harness_path()
    line = {"impl": "reference", "metric": "particle-iteration updates/sec", "unit": "particle-iteration updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}
    if os.path.exists(binp):
        dims = (20, 14, 20)
        i, j, k = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
        pos = np.stack([-0.95 + 0.1 * i, 0.05 + 0.1 * j, -0.95 + 0.1 * k], axis=-1).reshape(-1, 3).astype(np.float64)
        pos += np.random.default_rng(1234).uniform(-0.001, 0.001, size=pos.shape)
        vel = np.zeros_like(pos); vel[:, 1] = -1.0
        tmp = os.path.join(ROOT, "gpurun_out"); os.makedirs(tmp, exist_ok=True)
        scene = os.path.join(tmp, "ref_scene.bin"); dump = os.path.join(tmp, "ref_dump.bin")
        H.write_bin_scene(scene, pos, vel, RHO0)
        total = args.steps + args.warmup
        subprocess.run([binp, "--bin", scene, "--steps", str(total), "--out", dump, "--quiet"], check=True)
        secs = [d["seconds"] for d in H.read_dump(dump)][args.warmup:]
        n = pos.shape[0]
        val = n * ITERATIONS * len(secs) / sum(secs)
        sample = f"unmodified reference (oracle/_ref/ref_harness, -O3), {dims[0]}x{dims[1]}x{dims[2]} block = {n} particles in the Cornell box, O(N^2) search, 1 thread"
        line.update({"value": val, "ms_per_step": 1e3 * sum(secs) / len(secs),
                     "config": {"workload": "bounded sample of C4 dam-break lattice: " + sample, "iterations": ITERATIONS, "particles": n},
                     "cpu_baseline": {"value": val, "unit": line["unit"], "cores": 1, "kind": "reference", "sample": sample}})
    else:
        cb = cpu_baseline_port()
        line.update({"value": cb["value"], "ms_per_step": None,
                     "config": {"workload": "bounded sample of C4 dam-break lattice: " + cb["sample"], "iterations": ITERATIONS},
                     "cpu_baseline": cb})
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    line["gpu_launches"] = 0
    # Same-algorithm-class comparator next to the O(N^2) reference: the oracle port (uniform grid + OpenMP, fp64) on ALL
    # host cores at >= 1M particles of the C4 lattice (the unmodified reference cannot run that size at all).
    if os.path.exists(binp) and not args.no_cpu_baseline:
        try:
            port = cpu_baseline_port(min_side=104)
            port["nproc"] = os.cpu_count()
            line["cpu_baseline_port"] = port
        except Exception as e:      # the reference line itself must survive
            line["cpu_baseline_port"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)


# ---- GPU arm ------------------------------------------------------------------------------------------

def slab_witness(api, slab, dist, rank, world, local_rank, iters, steps=3):
    """Correctness witness carried by every N>1 line: a small tank (96*N x 20 x 20 particles, jittered, with x-motion so
    that particles migrate between slabs) stepped `steps` times as N slabs with the SAME transport the timed region uses,
    and as one handle on rank 0; positions, velocities, densities and neighbour digests must be bit-equal."""
    nx, ny, nz = 96 * world, 20, 20
    i, j, k = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), np.arange(nz, dtype=np.float64), indexing="ij")
    pos = np.stack([0.1 + 0.1 * i, 0.1 + 0.1 * j, 0.1 + 0.1 * k], axis=-1).reshape(-1, 3)
    rng = np.random.default_rng(77)
    pos += rng.uniform(-0.001, 0.001, size=pos.shape)
    vel = rng.normal(0.0, 0.3, size=pos.shape); vel[:, 1] -= 1.0; vel[:, 0] += 1.5 * np.sin(2.0 * pos[:, 0])
    box_max = (0.1 * nx + 0.4, 0.1 * ny + 2.0, 0.1 * nz + 0.3)
    prm = api.default_params(rest_density=RHO0, iterations=iters, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
    n = pos.shape[0]; per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    s = slab.SlabSolver(prm, rank, world, device=local_rank)
    s.upload_local(pos[lo:hi], vel[lo:hi], id_offset=lo)
    s.step(steps); s.sync()
    P, V, R, _ids, d, c = s.gather_all()
    out = {"particles": n, "steps": steps, "transport": s.transport, "n_conserved": bool(P.shape[0] == n)}
    if rank == 0:
        g = api.Solver(prm, device=local_rank)
        g.upload(pos, vel); g.step(steps)
        Pg, Vg, Rg = g.download(); dg, cg = g.neighbor_digest()
        same = P.shape[0] == n and all(np.array_equal(a, b) for a, b in ((P, Pg), (V, Vg), (R, Rg), (d, dg), (c, cg)))
        out["slab_equals_single"] = bool(same)
        if not same and P.shape[0] == n:
            out["max_abs_dpos"] = float(np.abs(P - Pg).max())
        del g
    del s
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, nargs=3, default=None, help="particles per GPU (nx ny nz); default 400 200 200")
    ap.add_argument("--iterations", type=int, default=ITERATIONS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tank", action="store_true", help="N=1 only: run the per-GPU wide-tank slab workload instead of C4 (weak-scaling reference point)")
    ap.add_argument("--settle", type=int, default=None, help="after the headline measurement let the dam break flow for K more steps and time the evolved "
                    "(disordered, ~75 neighbours) state as the extra key \"evolved\"; default 150 at N=1, 0 (off) at N>1")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # NCCL / driver banners must not pollute stdout: rank 0 prints exactly ONE JSON line there
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from fluid_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nx, ny, nz = args.block or BLOCK_PER_GPU
    iters = args.iterations
    hbm_peak, peak_src = measured_peaks()
    witness = None
    if world > 1:
        from fluid_b200 import slab as _slab
        witness = slab_witness(api, _slab, dist, rank, world, local_rank, iters)

    if world == 1 and not args.tank:
        box_max = (max(120.0, 0.3 * nx), 30.0, SPACING * nz + 0.1)
        params = api.default_params(rest_density=RHO0, iterations=iters, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
        solver = api.Solver(params, device=local_rank)
        pos, vel = block_f32(nx, ny, nz)
        n_local = pos.shape[0]
        workload = f"C4 dam-break block {nx}x{ny}x{nz} = {n_local} particles, spacing 0.1, H 0.3, box {box_max}, vorticity+XSPH on"
        step_fn = lambda k: solver.step(k, sync=False)
        sync_fn = solver.sync
        solver.upload(pos, vel)
    else:
        from fluid_b200 import slab
        box_max = (SPACING * nx * world + 0.1, 30.0, SPACING * nz + 0.1)
        params = api.default_params(rest_density=RHO0, iterations=iters, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2])
        pos, vel = block_f32(nx * world, ny, nz, chunk=rank, nchunks=world)
        n_local = pos.shape[0]
        solver = slab.SlabSolver(params, rank, world, device=local_rank)
        solver.upload_local(pos, vel, id_offset=rank * n_local)
        workload = f"C5-style wide tank {nx * world}x{ny}x{nz} = {n_local * world} particles in {world} x-slabs, box {box_max}, halo exchange + migration by peer stores over NVLink (transport {solver.transport})"
        step_fn = lambda k: solver.step(k)
        sync_fn = solver.sync

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up
    step_fn(args.warmup); sync_fn(); barrier()
    slab_mode = world > 1 or args.tank
    base = solver.solver if slab_mode else solver
    base.profile_enable(True)
    launches0 = base.launch_count()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    t0 = time.perf_counter()
    step_fn(args.steps); sync_fn()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = solver.last_ms() if slab_mode else solver.stats()[2]      # CUDA events on the solver's stream
    barrier()
    clocks = sampler.stop()
    launches = base.launch_count() - launches0
    prof = base.profile(); base.profile_enable(False)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    n_total = n_local * world
    ms_per_step = dev_ms / args.steps
    value = n_total * iters * args.steps / (dev_ms * 1e-3)

    # per-kernel table from CUDA events recorded on the launching stream during the timed region
    kernels = {}
    alg = {"lambda": BYTES_LAMBDA, "delta_collide": BYTES_DELTA}
    for name, (ms, cnt) in prof.items():
        if cnt:
            e = {"launches": cnt, "ms_per_launch": ms / cnt, "share": ms / max(dev_ms, 1e-9)}
            if name in alg:
                gbs = n_local * alg[name] / (ms / cnt * 1e-3) / 1e9
                e.update({"algorithmic_bytes_per_particle": alg[name], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
            kernels[name] = e
    dom = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches"])
    # N > 1: what every rank spent per step in its own kernels and in the exchange-point kernels (signal + bounded wait for the
    # neighbours), and its SM clock: the slabs run in lock step, so the slowest GPU of the chain sets the pace for all
    per_rank = None
    if world > 1:
        work = sum(ms for name, (ms, cnt) in prof.items() if name != "slab") / args.steps
        wait = prof.get("slab", (0.0, 0))[0] / args.steps
        mine = torch.tensor([work, wait, float(clocks.get("sm_mhz") or 0.0), float(clocks.get("power_w_max") or 0.0)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernels_ms_per_step": [round(float(a[0]), 3) for a in allr], "exchange_points_ms_per_step": [round(float(a[1]), 3) for a in allr],
                    "sm_mhz": [float(a[2]) for a in allr], "power_w_max": [float(a[3]) for a in allr]}
    traffic = None; ncu_note = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("particles") == n_local and dom in tj.get("kernels", {}):
            traffic = tj["kernels"][dom]["dram_bytes_per_launch"]
            ncu_note = {k: v for k, v in tj["kernels"][dom].items() if k != "dram_bytes_per_launch"}   # what ncu says binds it
            ncu_note["source"] = tj.get("source")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src, "ncu": ncu_note,
                "algorithmic_bytes_per_launch": n_local * alg[dom], "avg_launch_ms": kernels[dom]["ms_per_launch"],
                "whole_step": {"algorithmic_bytes_per_particle": BYTES_FIXED + (BYTES_LAMBDA + BYTES_DELTA) * iters,
                               "achieved": n_local * (BYTES_FIXED + (BYTES_LAMBDA + BYTES_DELTA) * iters) / (ms_per_step * 1e-3) / 1e9,
                               "frac": n_local * (BYTES_FIXED + (BYTES_LAMBDA + BYTES_DELTA) * iters) / (ms_per_step * 1e-3) / 1e9 / hbm_peak}}

    # State of the neighbour lists the headline was measured on, and the same measurement on an EVOLVED state: the
    # headline state is a few-step-old compressed lattice (~107 neighbours, very coherent gathers); a flowing dam
    # break has ~75 neighbours in disordered positions.  pairs/s = pair evaluations of the solver passes
    # (2 passes x I iterations x sum of list lengths) per second.
    def mean_neighbours():
        d, c = (solver.neighbor_digest() if slab_mode else base.neighbor_digest())
        t = torch.tensor([float(c.sum(dtype=np.float64)), float(len(c))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return float(t[0] / t[1])
    nb_head = mean_neighbours()
    settle = args.settle if args.settle is not None else (150 if world == 1 and not args.tank else 0)
    evolved = None
    if settle > 0:
        step_fn(settle); sync_fn(); barrier()
        base.profile_enable(True)
        k3 = max(3, min(args.steps, 10))
        step_fn(k3); sync_fn(); torch.cuda.synchronize()
        ev_ms = solver.last_ms() if slab_mode else solver.stats()[2]
        prof_e = base.profile(); base.profile_enable(False)
        t = torch.tensor([ev_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ev_ms = float(t[0]); nb_e = mean_neighbours()
        evolved = {"settle_steps": settle, "steps": k3, "ms_per_step": ev_ms / k3, "value": n_total * iters * k3 / (ev_ms * 1e-3),
                   "mean_neighbours": nb_e, "pairs_per_s": 2.0 * n_total * nb_e * iters * k3 / (ev_ms * 1e-3),
                   "kernels_ms_per_launch": {k: ms / cnt for k, (ms, cnt) in prof_e.items() if cnt}}

    # e2e: host buffers in and out EVERY step through the C ABI (pbf_upload -> pbf_step -> pbf_download)
    e2e = None
    if not args.no_e2e:
        k2 = max(2, min(args.steps, 5))
        if not slab_mode:
            P = np.empty((n_local, 3)); V = np.empty((n_local, 3)); R = np.empty(n_local)
            solver.pin(P, V, R)          # the host arrays a caller reuses every step, page-locked once (pbf_host_register)
            solver.download_into(P, V, R)
            solver.set_readback(P, V, R)   # each step streams its result into P, V, R (complete after sync)
            barrier(); t0 = time.perf_counter()
            for _ in range(k2):
                solver.upload(P, V); solver.step(1, sync=True)
            torch.cuda.synchronize(); e_sync_ms = (time.perf_counter() - t0) * 1e3
            # Throughput form of the same loop: TWO handles (two independent batches of the same workload, each with its
            # own page-locked host buffers and its own stream) alternate, so that one batch's host<->device copies run
            # while the other batch computes.  Every step of either batch still uploads its inputs from the host and
            # streams its results back to the host inside the timed region.
            solver2 = api.Solver(params, device=local_rank)
            P2 = P.copy(); V2 = V.copy(); R2 = np.empty(n_local)
            solver2.pin(P2, V2, R2); solver2.upload(P2, V2); solver2.set_readback(P2, V2, R2)
            solver2.step(1, sync=True)                                  # allocate / warm up the second handle
            barrier(); t0 = time.perf_counter()
            solver.upload(P, V); solver.step(1, sync=False)
            for it in range(k2):
                solver2.upload(P2, V2); solver2.step(1, sync=False)     # copies of batch 2 overlap the step of batch 1
                solver.sync()                                           # batch 1: results are in P, V, R
                if it + 1 < k2:
                    solver.upload(P, V); solver.step(1, sync=False)     # copies of batch 1 overlap the step of batch 2
                solver2.sync()
            torch.cuda.synchronize(); e_ms = (time.perf_counter() - t0) * 1e3 / 2.0    # 2 * k2 steps were run: time per k2 steps
            del solver2
        else:
            capn = int(solver.particle_cap)
            P = np.empty((capn, 3)); V = np.empty((capn, 3)); R = np.empty(capn); I = np.empty(capn, dtype=np.uint32)
            solver.pin(P, V, R, I)
            nl = solver.download_local_into(P, V, R, I)
            barrier(); t0 = time.perf_counter()
            for _ in range(k2):
                solver.upload_local(P[:nl], V[:nl], ids=I[:nl]); solver.step(1); nl = solver.download_local_into(P, V, R, I)
            torch.cuda.synchronize(); e_ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t[0])
        # value = ONE simulation, one handle: upload -> step -> results back on the host, step after step (a step depends on
        # the previous one, so nothing of the next step can overlap).  The two-handle figure is a different workload
        # shape (two independent simulations sharing the GPU) and is reported beside it, not as the headline.
        h2d = n_total * (6 * 8 + (4 if world > 1 else 0)); d2h = n_total * (7 * 8 + (4 if world > 1 else 0))
        one_ms = e_sync_ms if not slab_mode else e_ms
        e2e = {"value": n_total * iters * k2 / (one_ms * 1e-3), "unit": "particle-iteration updates/s", "steps": k2,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": one_ms / k2,
               "host_gbs": (h2d + d2h) / (max(one_ms / k2 - ms_per_step, 1e-3) * 1e-3) / 1e9,
               "note": "ONE simulation through one handle: host fp64 AoS buffers (pos, vel) uploaded (pbf_upload) and (pos, vel, density) read back EVERY "
                       "step (streaming read-back behind the finalize kernels); page-locked caller buffers, fp64 on the wire, fp64<->fp32 on the device; "
                       "wall clock; host_gbs = copied bytes / (e2e time - device step time), the rate of the copies that are NOT hidden"}
        if not slab_mode:
            e2e["pipelined_two_batches_value"] = n_total * iters * k2 / (e_ms * 1e-3)
            e2e["pipelined_two_batches_ms_per_step"] = e_ms / k2
            e2e["pipelined_note"] = "two independent simulations (two handles, own page-locked buffers and streams) alternating, one batch's copies behind the other's step: throughput of a different workload shape"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.tank:
        cpu = cpu_baseline_port()

    if rank == 0:
        line = {"metric": "particle-iteration updates/sec", "value": value, "unit": "particle-iteration updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "particles": n_total, "particles_per_gpu": n_local, "iterations": iters,
                           "l2_policy": "inputs larger than L2 (256 MB per float4 array vs 126 MB L2); no flush needed",
                           "parallelism": "single GPU" if world == 1 else (f"{world} x-slabs, one process per GPU, halo exchange by peer stores (CUDA IPC) + flag hand-overs"
                                                                            if getattr(solver, "transport", "") == "p2p" else
                                                                            f"{world} x-slabs, one process per GPU, halo exchange by NCCL send/recv (transport {getattr(solver, 'transport', '?')}: peer mapping unavailable)")},
                "wall_ms_per_step": wall_ms / args.steps, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "mean_neighbours": nb_head, "pairs_per_s": 2.0 * n_total * nb_head * iters / (ms_per_step * 1e-3), "evolved": evolved,
                "gpu_launches": int(launches), "clocks": clocks, "kernels": kernels}
        if witness is not None:
            line["slab_equals_single"] = witness.get("slab_equals_single"); line["n_conserved"] = witness["n_conserved"]; line["witness"] = witness
        if per_rank is not None:
            line["per_rank"] = per_rank
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
