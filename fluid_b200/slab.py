"""x-slab decomposition of the PBF step over the GPUs of one box: one process per GPU (the driver of the
single-process form is pbf_create_multi, include/pbf_b200_multi.h); every kernel is in libpbf_b200.so, C ABI
include/pbf_b200_slab.h.  Two transports between x-adjacent ranks:
  "p2p"   (default with one GPU per rank) peer mode of the library: CUDA IPC maps the neighbours' buffers once, then
          migration / ghost messages and the per-iteration boundary values are stored straight into the neighbours'
          memory by the kernels that produce them and hand-overs are flag words; a step is ONE library call with no
          host synchronisation, torch.distributed only carries the IPC handles and the re-balancing histogram;
  "nccl"  the host-driven protocol below with torch.distributed send/recv of device ranges (and "staged": the same
          through host copies, for gloo).

Protocol per step (the reference is single-process, particles.cpp:250-297):
  predict -> exchange emigrants -> absorb + pack ghost layers -> exchange ghosts -> sort + neighbour
  lists -> I x (lambda, refresh ghost (x*,lambda); delta-p, refresh ghost x*) -> velocity,
  vorticity/XSPH, refresh ghost |omega|, confinement + commit.
Because the global cell grid and the in-cell order (global particle id) are the same on every
rank, the owner's boundary column and the neighbour's ghost column are ordered identically, so the
per-iteration refreshes are contiguous device-to-device sends of 16 B per ghost with no packing,
and an N-slab run reproduces the single-GPU run bit for bit.
"""
import ctypes as C

import numpy as np

from . import api

(BUF_MIG_SEND_L, BUF_MIG_SEND_R, BUF_MIG_RECV_L, BUF_MIG_RECV_R, BUF_GHOST_SEND_L, BUF_GHOST_SEND_R,
 BUF_GHOST_RECV_L, BUF_GHOST_RECV_R, BUF_XS_A, BUF_XS_B, BUF_OMEGA, BUF_XS_W) = range(12)
PH_LAMBDA_FIRST, PH_LAMBDA, PH_DELTA, PH_VELOCITY, PH_VORTICITY, PH_CONFINE = range(6)
PART_ALL, PART_BOUNDARY, PART_INTERIOR = range(3)


def _bind(lib):
    if getattr(lib, "_slab_bound", False):
        return lib
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    sig = {
        "pbf_grid_dims": (i32, [C.POINTER(api.PbfParams), C.POINTER(C.c_int * 3)]),
        "pbf_cell_columns": (i32, [C.POINTER(api.PbfParams), sz, vp, vp]),
        "pbf_set_stream": (i32, [vp, vp]),
        "pbf_slab_configure": (i32, [vp, i32, i32, i32, i32, sz, sz]),
        "pbf_slab_configure_ex": (i32, [vp, i32, i32, i32, i32, sz, sz, i32]),
        "pbf_slab_set_columns": (i32, [vp, i32, i32, i32, i32]),
        "pbf_slab_columns": (i32, [vp, C.POINTER(C.c_int * 4)]),
        "pbf_slab_column_histogram": (i32, [vp, vp, sz, i32, C.POINTER(C.c_longlong)]),
        "pbf_slab_p2p_blob_size": (sz, []),
        "pbf_slab_p2p_export": (i32, [vp, vp]),
        "pbf_slab_p2p_connect_ipc": (i32, [vp, vp, vp]),
        "pbf_slab_set_wait_timeout": (i32, [vp, C.c_double]),
        "pbf_slab_p2p_disconnect": (i32, [vp]),
        "pbf_slab_step_p2p": (i32, [vp, i32]),
        "pbf_slab_set_histogram_interval": (i32, [vp, i32]),
        "pbf_slab_refresh_ranges": (i32, [vp, C.POINTER(C.c_uint32 * 5)]),
        "pbf_partition_columns": (i32, [vp, i32, i32, vp]),
        "pbf_plan_rebalance": (i32, [vp, i32, i32, vp, C.c_uint64, C.c_double, vp, C.POINTER(C.c_double)]),
        "pbf_slab_upload": (i32, [vp, sz, vp, vp, vp]),
        "pbf_slab_download": (i32, [vp, sz, vp, vp, vp, vp, C.POINTER(sz)]),
        "pbf_slab_neighbor_digest": (i32, [vp, sz, vp, vp]),
        "pbf_slab_phase_predict": (i32, [vp]),
        "pbf_slab_phase_migrate": (i32, [vp]),
        "pbf_slab_phase_sort": (i32, [vp, C.POINTER(C.c_uint32 * 5)]),
        "pbf_slab_phase": (i32, [vp, i32]),
        "pbf_slab_phase_part": (i32, [vp, i32, i32]),
        "pbf_slab_stats": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
        "pbf_slab_buffer": (vp, [vp, i32, C.POINTER(sz)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._slab_bound = True
    return lib


# ---- pure host logic (unit-tested on CPU, tests/test_slab_host.py) -----------------------------------

def partition_columns(hist, world):
    """Split cell columns 0..len(hist) into `world` contiguous slabs with (nearly) equal particle
    counts.  Returns world+1 boundaries; every slab gets at least one column."""
    hist = np.asarray(hist, dtype=np.int64)
    ncol = len(hist)
    if world > ncol:
        raise ValueError(f"{world} slabs need at least {world} cell columns, the box has {ncol}")
    csum = np.concatenate([[0], np.cumsum(hist)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        # nearest boundary to the target count
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, ncol)] - target):
            b -= 1
        b = max(b, bounds[-1] + 1)
        b = min(b, ncol - (world - r))
        bounds.append(b)
    bounds.append(ncol)
    return bounds


def neighbours_of(rank, world):
    return (rank - 1 if rank > 0 else None), (rank + 1 if rank < world - 1 else None)


def exchange(dist, ops_spec, group=None, staged=False):
    """ops_spec: list of (kind, tensor, peer) with kind in {'send','recv'}; peers may be None (skipped).
    NCCL: one batched group, ordered on the current CUDA stream (device-to-device over NVLink).
    staged=True (gloo backend: CPU tests, or several ranks sharing one GPU): the same transfers
    through host copies."""
    spec = [(k, t, p) for k, t, p in ops_spec if p is not None and t.numel() > 0]
    if not spec:
        return
    if not staged:
        ops = [dist.P2POp(dist.isend if k == "send" else dist.irecv, t, p, group) for k, t, p in spec]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return
    import torch
    if any(t.is_cuda for _, t, _ in spec):
        torch.cuda.current_stream().synchronize()
    work, recvs = [], []
    for k, t, p in spec:
        if k == "send":
            work.append(dist.isend(t.detach().cpu().contiguous(), p, group))
        else:
            tmp = torch.empty(t.shape, dtype=t.dtype)
            recvs.append((t, tmp))
            work.append(dist.irecv(tmp, p, group))
    for w in work:
        w.wait()
    for t, tmp in recvs:
        t.copy_(tmp)


class _DevMem:
    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class SlabSolver:
    """One rank of the slab-decomposed solver."""

    def __init__(self, params, rank, world, device=0, halo_factor=4.0, staged=None, overlap=None, transport=None,
                 rebalance_every=8, rebalance_threshold=1.05, cap_factor=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device = rank, world, device
        self.params = params
        self.lib = _bind(api.load_library())
        self.solver = api.Solver(params, device)          # plain handle; configured as a slab at upload
        self.h = self.solver.h
        self.halo_factor = halo_factor
        self.left, self.right = neighbours_of(rank, world)
        dims = (C.c_int * 3)()
        self._ck(self.lib.pbf_grid_dims(C.byref(params), C.byref(dims)))
        self.gdims = tuple(dims)
        self.stream = torch.cuda.Stream(device=device)
        self._ck(self.lib.pbf_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)))
        self.bounds = None
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._last_ms = 0.0
        self.iterations = params.iterations
        # gloo cannot move device memory: stage through the host (tests; ranks sharing one GPU)
        self.staged = (world > 1 and dist.get_backend() != "nccl") if staged is None else staged
        # Optional (PBF_SLAB_OVERLAP=1): overlap each pass's halo exchange with the interior part of the
        # same pass on a second stream.  Measured at 2 x 16M particles: 88.1 ms/step overlapped vs
        # 87.5 ms sequential — the exchanges (~2 MB, NVLink) are not what the step waits for — so off by default.
        import os
        want = os.environ.get("PBF_SLAB_OVERLAP", "0") == "1" if overlap is None else overlap
        self.overlap = bool(want) and world > 1 and not self.staged
        self.comm_stream = torch.cuda.Stream(device=device) if self.overlap else None
        # transport: "p2p" = the library's peer mode (CUDA IPC between the ranks' processes)
        t = transport or os.environ.get("PBF_SLAB_TRANSPORT")
        if t is None:
            t = "p2p" if (world == 1 or not self.staged) and not self.overlap else ("staged" if self.staged else "nccl")
        assert t in ("p2p", "nccl", "staged"), t
        self.transport = t
        if t == "staged":
            self.staged = True
        self.rebalance_every, self.rebalance_threshold = int(rebalance_every), float(rebalance_threshold)
        self.cap_factor = cap_factor if cap_factor is not None else (1.30 if rebalance_every else 1.15)
        self.steps_done = 0
        self.last_rebalance_at = 0
        self.n_rebalances = 0

    def _ck(self, rc):
        if rc != api.PBF_OK:
            raise api.PbfError(rc, self.lib.pbf_last_error(self.h).decode())

    # -- setup ---------------------------------------------------------------------------------------
    def set_obstacle_spheres(self, spheres):
        """Obstacle spheres (rows cx, cy, cz, r): global scene data, call with the same list on every rank."""
        self.solver.set_obstacle_spheres(spheres)

    def set_obstacle_triangles(self, tris):
        """Obstacle triangles (rows of 18): global scene data, call with the same list on every rank."""
        self.solver.set_obstacle_triangles(tris)

    def columns_of(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        col = np.empty(pos.shape[0], dtype=np.int32)
        self._ck(self.lib.pbf_cell_columns(C.byref(self.params), pos.shape[0], pos.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p)))
        return col

    def plan(self, pos_local):
        """Balanced column partition from the global column histogram (all ranks)."""
        torch, dist = self.torch, self.dist
        hist = np.bincount(self.columns_of(pos_local), minlength=self.gdims[0]).astype(np.int64)
        t = torch.from_numpy(hist)
        if self.world > 1:
            t = t.to(f"cuda:{self.device}") if dist.get_backend() == "nccl" else t
            dist.all_reduce(t)
        hist = t.cpu().numpy()
        return partition_columns(hist, self.world), hist

    def upload_local(self, pos, vel, id_offset=0, ids=None, plan=None):
        """pos/vel: this rank's share of the particles (any particle may sit one slab off: the first
        step migrates it).  Global ids default to id_offset + arange."""
        pos = np.ascontiguousarray(pos, dtype=np.float64); vel = np.ascontiguousarray(vel, dtype=np.float64)
        n = pos.shape[0]
        if ids is None:
            ids = np.arange(id_offset, id_offset + n, dtype=np.uint32)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        if self.bounds is None:
            bnds, hist = plan if plan is not None else self.plan(pos)
            self.col_bounds = bnds
            lo, hi = bnds[self.rank], bnds[self.rank + 1]
            left_cols = bnds[self.rank] - bnds[self.rank - 1] if self.left is not None else 0
            right_cols = bnds[self.rank + 2] - bnds[self.rank + 1] if self.right is not None else 0
            per_col = max(int(hist.max()), 1)           # global, so message sizes agree on both ends
            self.halo_cap = int(max(4096, self.halo_factor * per_col))
            owned_est = int(hist[lo:hi].sum())
            mean_owned = int(hist.sum()) // self.world + 1
            self.particle_cap = int(max(n, owned_est, mean_owned) * self.cap_factor + 2 * self.halo_cap + 1024)
            ncol = self.gdims[0]
            cells_per_col = self.gdims[1] * self.gdims[2]
            max_cols = ncol if cells_per_col * ncol * 8.0 <= 2e9 else min(ncol, 4 * (hi - lo) + 16)
            self._ck(self.lib.pbf_slab_configure_ex(self.h, lo, hi, left_cols, right_cols, self.particle_cap, self.halo_cap, max_cols))
            self._map_buffers()
            if self.rebalance_every:
                self._ck(self.lib.pbf_slab_set_histogram_interval(self.h, self.rebalance_every))   # recorded on named steps: a deterministic schedule
            if self.transport == "p2p":
                self._connect_p2p()
            self.bounds = (0, 0, 0, 0, 0)
        self._ck(self.lib.pbf_slab_upload(self.h, n, pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p)))
        self.solver.n = n

    def _connect_p2p(self):
        """Exchange CUDA IPC handles of the buffers the x-neighbours write into; from here on the step needs no transport.
        If ANY rank cannot map its neighbours (no peer access between two GPUs, IPC refused by the container) every rank
        falls back to the host-driven exchange over the process group (NCCL send/recv), loudly."""
        dist, torch = self.dist, self.torch
        nb = self.lib.pbf_slab_p2p_blob_size()
        blob = C.create_string_buffer(nb)
        ok, why = 1, ""
        if self.world > 1 and self.lib.pbf_slab_p2p_export(self.h, blob) != api.PBF_OK:     # a lone slab has nobody to export to
            ok, why = 0, self.lib.pbf_last_error(self.h).decode()
        blobs = [blob.raw if ok else None] * self.world
        if self.world > 1:
            dist.all_gather_object(blobs, blob.raw if ok else None)
        if ok and all(b is not None for b in blobs):
            left = C.create_string_buffer(blobs[self.left], nb) if self.left is not None else None
            right = C.create_string_buffer(blobs[self.right], nb) if self.right is not None else None
            if self.lib.pbf_slab_p2p_connect_ipc(self.h, left, right) != api.PBF_OK:
                ok, why = 0, self.lib.pbf_last_error(self.h).decode()
        else:
            ok = 0
        if self.world > 1:
            t = torch.tensor([ok], dtype=torch.int32)
            t = t.to(f"cuda:{self.device}") if dist.get_backend() == "nccl" else t
            dist.all_reduce(t, op=dist.ReduceOp.MIN)      # also the barrier: nobody steps before every neighbour is mapped
            all_ok = int(t.item())
        else:
            all_ok = ok
        if not all_ok:
            import sys
            self.p2p_fallback_reason = why or "another rank could not map its neighbours"
            if why:
                print(f"[fluid_b200.slab] rank {self.rank}: peer mode unavailable ({why}); falling back to the process-group exchange", file=sys.stderr, flush=True)
            self._ck(self.lib.pbf_slab_p2p_disconnect(self.h))
            self.transport = "staged" if self.staged else "nccl"

    def _map_buffers(self):
        torch = self.torch
        self.buf = {}
        for which in range(12):
            nbytes = C.c_size_t()
            ptr = self.lib.pbf_slab_buffer(self.h, which, C.byref(nbytes))
            t = torch.as_tensor(_DevMem(ptr, nbytes.value // 4), device=f"cuda:{self.device}")
            self.buf[which] = t.view(-1, 4)

    # -- one step --------------------------------------------------------------------------------------
    def _halo(self, which):
        """Refresh the ghost ranges of array `which` from the owners' boundary columns."""
        b0, b1, b2, b3, n = self.bounds
        a = self.buf[which]
        exchange(self.dist, [("send", a[b0:b1], self.left), ("recv", a[0:b0], self.left),
                             ("send", a[b2:b3], self.right), ("recv", a[b3:n], self.right)], staged=self.staged)

    def _rebalance(self):
        """Between two steps: move the slab boundaries towards equal counts when the fluid has flowed (every rank takes
        the same decision from the same all-reduced histogram; the next predict pass migrates the particles)."""
        torch, dist, lib = self.torch, self.dist, self.lib
        if self.world < 2 or not self.rebalance_every or self.steps_done == 0 or self.steps_done % self.rebalance_every:
            return
        ncol = self.gdims[0]
        hist = np.zeros(ncol, dtype=np.uint32); at = C.c_longlong(-1)
        rc = lib.pbf_slab_column_histogram(self.h, hist.ctypes.data_as(C.c_void_p), ncol, 1, C.byref(at))
        lo, hi = self.col_bounds[self.rank], self.col_bounds[self.rank + 1]
        own = np.zeros(ncol + 1, dtype=np.int64); own[lo:hi] = hist[lo:hi]
        own[ncol] = 1 if rc == api.PBF_OK and at.value >= self.last_rebalance_at else 0     # taken under the current plan? (the all-reduce below is collective either way)
        t = torch.from_numpy(own)
        t = t.to(f"cuda:{self.device}") if dist.get_backend() == "nccl" else t
        dist.all_reduce(t)
        own = t.cpu().numpy()
        if own[ncol] != self.world:
            return
        g = np.ascontiguousarray(own[:ncol], dtype=np.uint64)
        old = np.ascontiguousarray(self.col_bounds, dtype=np.int32); new = np.zeros_like(old); imb = C.c_double()
        ch = lib.pbf_plan_rebalance(g.ctypes.data_as(C.c_void_p), ncol, self.world, old.ctypes.data_as(C.c_void_p), int(0.4 * self.halo_cap),
                                    self.rebalance_threshold, new.ctypes.data_as(C.c_void_p), C.byref(imb))
        self.imbalance = imb.value
        if ch != 1:
            return
        b = [int(v) for v in new]
        r = self.rank
        self._ck(lib.pbf_slab_set_columns(self.h, b[r], b[r + 1], b[r] - b[r - 1] if r > 0 else 0, b[r + 2] - b[r + 1] if r + 1 < self.world else 0))
        self.col_bounds = b
        self.last_rebalance_at = self.steps_done
        self.n_rebalances += 1

    def _step_once(self):
        lib, h, B = self.lib, self.h, self.buf
        self._rebalance()
        self.steps_done += 1
        if self.transport == "p2p":
            self._ck(lib.pbf_slab_step_p2p(h, 1))
            return
        self._ck(lib.pbf_slab_phase_predict(h))
        exchange(self.dist, [("send", B[BUF_MIG_SEND_L], self.left), ("recv", B[BUF_MIG_RECV_L], self.left),
                             ("send", B[BUF_MIG_SEND_R], self.right), ("recv", B[BUF_MIG_RECV_R], self.right)], staged=self.staged)
        self._ck(lib.pbf_slab_phase_migrate(h))
        exchange(self.dist, [("send", B[BUF_GHOST_SEND_L], self.left), ("recv", B[BUF_GHOST_RECV_L], self.left),
                             ("send", B[BUF_GHOST_SEND_R], self.right), ("recv", B[BUF_GHOST_RECV_R], self.right)], staged=self.staged)
        out = (C.c_uint32 * 5)()
        self._ck(lib.pbf_slab_phase_sort(h, C.byref(out)))
        self.bounds = tuple(int(v) for v in out)
        if not self.overlap:
            for it in range(self.iterations):
                self._ck(lib.pbf_slab_phase(h, PH_LAMBDA_FIRST if it == 0 else PH_LAMBDA))
                self._halo(BUF_XS_B)
                self._ck(lib.pbf_slab_phase(h, PH_DELTA))
                self._halo(BUF_XS_A)
            self._ck(lib.pbf_slab_phase(h, PH_VELOCITY))
            self._ck(lib.pbf_slab_phase(h, PH_VORTICITY))
            self._halo(BUF_XS_W)          # ghost (x*, |omega|): the confinement pass gathers one float4 per pair
            self._ck(lib.pbf_slab_phase(h, PH_CONFINE))
            return
        # overlapped: boundary columns first, their exchange runs on the comm stream while the
        # interior of the same pass computes; the next pass waits for the exchange
        for it in range(self.iterations):
            self._pass_overlapped(PH_LAMBDA_FIRST if it == 0 else PH_LAMBDA, BUF_XS_B)
            self._pass_overlapped(PH_DELTA, BUF_XS_A)
        self._ck(lib.pbf_slab_phase(h, PH_VELOCITY))
        self._pass_overlapped(PH_VORTICITY, BUF_XS_W)
        self._ck(lib.pbf_slab_phase(h, PH_CONFINE))

    def _pass_overlapped(self, phase, which):
        torch = self.torch
        self._ck(self.lib.pbf_slab_phase_part(self.h, phase, PART_BOUNDARY))
        ready = torch.cuda.Event(); ready.record(self.stream)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            self._halo(which)
            done = torch.cuda.Event(); done.record(self.comm_stream)
        self._ck(self.lib.pbf_slab_phase_part(self.h, phase, PART_INTERIOR))
        self.stream.wait_event(done)

    def step(self, n_steps=1):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self._ev[0].record(self.stream)
            for _ in range(n_steps):
                self._step_once()
            self._ev[1].record(self.stream)
        self._timed = True

    def sync(self):
        self.stream.synchronize()
        self.solver.sync()
        if self.transport == "p2p" and self.bounds is not None:
            out = (C.c_uint32 * 5)()
            self._ck(self.lib.pbf_slab_refresh_ranges(self.h, C.byref(out)))     # the ranges stayed on the device during the step
            self.bounds = tuple(int(v) for v in out)
        if getattr(self, "_timed", False):
            self._last_ms = self._ev[0].elapsed_time(self._ev[1]); self._timed = False

    def last_ms(self):
        self.sync()
        return self._last_ms

    # -- results ---------------------------------------------------------------------------------------
    def n_owned(self):
        return self.bounds[3] - self.bounds[0] if self.bounds and self.bounds[4] else self.solver.n

    def download_local(self, with_ids=False):
        self.sync()
        cap = max(self.n_owned(), self.solver.n, 1)
        P = np.empty((cap, 3)); V = np.empty((cap, 3)); R = np.empty(cap); I = np.empty(cap, dtype=np.uint32)
        n = C.c_size_t()
        self._ck(self.lib.pbf_slab_download(self.h, cap, P.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p),
                                            R.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p), C.byref(n)))
        n = n.value
        self._ids = I[:n].copy()
        return (P[:n], V[:n], R[:n], I[:n]) if with_ids else (P[:n], V[:n], R[:n])

    def local_ids(self):
        return self._ids

    def pin(self, *arrays):
        """Page-lock caller arrays reused for upload_local / download_local_into (pbf_host_register)."""
        self.solver.pin(*arrays)

    def download_local_into(self, P, V, R, I):
        """Owned particles into caller-provided (ideally pinned) arrays of capacity >= n_owned; returns n."""
        self.sync()
        n = C.c_size_t()
        self._ck(self.lib.pbf_slab_download(self.h, P.shape[0], P.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p),
                                            R.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p), C.byref(n)))
        return n.value

    def neighbor_digest(self):
        self.sync()
        cap = max(self.n_owned(), 1)
        d = np.empty(cap, dtype=np.uint64); c = np.empty(cap, dtype=np.uint32)
        self._ck(self.lib.pbf_slab_neighbor_digest(self.h, cap, d.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
        return d[:self.n_owned()], c[:self.n_owned()]

    def stats(self):
        """Global avg density after the first lambda pass / after finalize (the reference's printout)."""
        torch, dist = self.torch, self.dist
        self.sync()
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        self._ck(self.lib.pbf_slab_stats(self.h, C.byref(a), C.byref(b), C.byref(n)))
        t = torch.tensor([a.value, b.value, float(n.value)], dtype=torch.float64)
        if self.world > 1:
            t = t.to(f"cuda:{self.device}") if dist.get_backend() == "nccl" else t
            dist.all_reduce(t)
        return float(t[0] / t[2]), float(t[1] / t[2])

    def gather_all(self):
        """All particles on every rank, ordered by global id: (pos, vel, rho, digest, count)."""
        torch, dist = self.torch, self.dist
        P, V, R, I = self.download_local(with_ids=True)
        d, c = self.neighbor_digest()
        local = (P, V, R, I, d, c)
        if self.world == 1:
            parts = [local]
        else:
            parts = [None] * self.world
            dist.all_gather_object(parts, local)
        P = np.concatenate([p[0] for p in parts]); V = np.concatenate([p[1] for p in parts]); R = np.concatenate([p[2] for p in parts])
        I = np.concatenate([p[3] for p in parts]); d = np.concatenate([p[4] for p in parts]); c = np.concatenate([p[5] for p in parts])
        o = np.argsort(I, kind="stable")
        return P[o], V[o], R[o], I[o], d[o], c[o]
