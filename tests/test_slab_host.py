"""Host-side logic of the slab decomposition on CPU: column partitioning and the halo-exchange
plumbing across two and three gloo ranks (world_size > 1, no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_partition_balanced_and_contiguous():
    from fluid_b200.slab import partition_columns
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        hist = np.concatenate([np.zeros(3, int), rng.integers(900, 1100, size=400), np.zeros(5, int)])
        b = partition_columns(hist, world)
        assert b[0] == 0 and b[-1] == len(hist) and len(b) == world + 1
        assert all(b[i] < b[i + 1] for i in range(world))
        counts = [hist[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(counts) - min(counts) <= 2 * hist.max()
    # more slabs than occupied columns still gives every slab a column
    b = partition_columns(np.array([0, 0, 50, 0, 0, 0]), 4)
    assert all(b[i] < b[i + 1] for i in range(4))
    with pytest.raises(ValueError):
        partition_columns(np.ones(3, int), 4)


def test_neighbours():
    from fluid_b200.slab import neighbours_of
    assert neighbours_of(0, 1) == (None, None)
    assert neighbours_of(0, 4) == (None, 1) and neighbours_of(2, 4) == (1, 3) and neighbours_of(3, 4) == (2, None)


def _ring_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fluid_b200.slab import exchange, neighbours_of
    left, right = neighbours_of(rank, world)
    ok = True
    for it in range(3):
        # a "sorted array" with ghost ranges at both ends: [ghostL | owned | ghostR]; boundary
        # columns have rank-dependent sizes, like real slabs
        nb_l, nb_r = 3 + rank, 4 + rank                       # my boundary-column sizes
        ng_l = (4 + rank - 1) if left is not None else 0      # = left's right-boundary size
        ng_r = (3 + rank + 1) if right is not None else 0     # = right's left-boundary size
        own = 20
        a = torch.zeros(ng_l + own + ng_r, 4)
        a[ng_l:ng_l + own] = rank * 1000 + it * 100 + torch.arange(own, dtype=torch.float32)[:, None]
        b0, b1, b2, b3, n = ng_l, ng_l + nb_l, ng_l + own - nb_r, ng_l + own, ng_l + own + ng_r
        exchange(dist, [("send", a[b0:b1], left), ("recv", a[0:b0], left), ("send", a[b2:b3], right), ("recv", a[b3:n], right)], staged=True)
        if left is not None:
            exp = (rank - 1) * 1000 + it * 100 + torch.arange(own - (4 + rank - 1), own, dtype=torch.float32)
            ok &= bool(torch.equal(a[0:b0, 0], exp))
        if right is not None:
            exp = (rank + 1) * 1000 + it * 100 + torch.arange(0, 3 + rank + 1, dtype=torch.float32)
            ok &= bool(torch.equal(a[b3:n, 0], exp))
    # fixed-size messages with a count header (migration / ghost messages)
    msg_out = torch.full((9, 4), float(rank)); msg_in_l = torch.zeros(9, 4); msg_in_r = torch.zeros(9, 4)
    exchange(dist, [("send", msg_out, left), ("recv", msg_in_l, left), ("send", msg_out, right), ("recv", msg_in_r, right)], staged=True)
    if left is not None:
        ok &= bool((msg_in_l == rank - 1).all())
    if right is not None:
        ok &= bool((msg_in_r == rank + 1).all())
    q.put((rank, ok))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29541), (3, 29542)])
def test_halo_exchange_plumbing_gloo(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
