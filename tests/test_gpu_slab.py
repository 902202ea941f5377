"""Slab decomposition == single GPU, bit for bit (GPU tests).  The 2-rank cases run as separate
processes under torch.distributed.run; with one visible GPU both ranks share it and the halos are
staged through the host (gloo), with two or more they use NCCL device-to-device."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "slab_worker.py")


def _run(world, extra, port, env=None):
    if world == 1:
        cmd = [sys.executable, WORKER] + extra
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), WORKER] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    lines = [l for l in r.stdout.splitlines() if l.startswith("SLAB_RESULT ")]
    assert r.returncode == 0 and lines, f"worker failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return json.loads(lines[-1][len("SLAB_RESULT "):])


def _check(res):
    assert res["ids_ok"] and res["finite"]
    assert res["digest_equal"], "neighbour sets differ between slab and single-GPU runs"
    assert res["pos_equal"] and res["vel_equal"] and res["rho_equal"], f"slab run is not bit-identical (max dpos {res['max_dpos']})"
    assert abs(res["avg_rho_slab"][0] - res["avg_rho_single"][0]) < 1e-6 * 700 and abs(res["avg_rho_slab"][1] - res["avg_rho_single"][1]) < 1e-6 * 700


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_slab_codepath_world1():
    """One rank, slab code path (phases, no neighbours) == pbf_step: host-driven phases and the peer-mode step."""
    _check(_run(1, ["--steps", "4", "--transport", "nccl"], 0))
    res = _run(1, ["--steps", "4", "--transport", "p2p"], 0)
    _check(res)
    assert res["transport"] == "p2p"


def test_two_slabs_shared_gpu_p2p_ipc():
    """Two PROCESSES, peer mode through CUDA IPC (here both on one GPU): messages and per-iteration boundary values are
    stored straight into the other process's buffers, hand-overs are flag words, no transport inside the step."""
    res = _run(2, ["--backend", "gloo", "--same-gpu", "--steps", "6", "--transport", "p2p"], 29618)
    _check(res)
    assert res["transport"] == "p2p"


def test_three_slabs_shared_gpu_p2p_ipc_rebalanced():
    """Three processes, peer mode, a block flowing along a long tank: the ranks re-balance from the all-reduced column
    histogram (same decision everywhere), the particles move through the ordinary migration messages."""
    res = _run(3, ["--backend", "gloo", "--same-gpu", "--steps", "120", "--transport", "p2p", "--flow", "--dims", "240", "10", "8",
                   "--rebalance-every", "2"], 29619)
    _check(res)
    assert res["n_rebalances"] >= 3
    assert max(res["owned"]) / (sum(res["owned"]) / 3) <= 1.08, res["owned"]


def test_two_slabs_shared_gpu_gloo():
    """Two ranks on ONE GPU, host-staged exchange: migration + ghosts + per-iteration refresh."""
    _check(_run(2, ["--backend", "gloo", "--same-gpu", "--steps", "6", "--transport", "staged"], 29611))


def test_two_slabs_with_obstacle_spheres():
    """Obstacle spheres are global scene data set on every rank; one straddles the slab boundary."""
    res = _run(2, ["--backend", "gloo", "--same-gpu", "--steps", "6", "--spheres", "--transport", "staged"], 29616)
    _check(res)
    assert -2e-6 <= res["min_sphere_gap"] <= 1e-3, res["min_sphere_gap"]     # particles rest on the spheres, none inside


def test_two_slabs_with_obstacle_mesh():
    """A 1520-triangle obstacle mesh (device BVH) across the slab boundary: every rank holds the whole hierarchy."""
    res = _run(2, ["--backend", "gloo", "--same-gpu", "--steps", "6", "--mesh", "--transport", "staged"], 29617)
    _check(res)
    assert res["inside_mesh"] == 0


def test_three_slabs_shared_gpu_gloo():
    """Three ranks: the middle slab has neighbours on both sides."""
    _check(_run(3, ["--backend", "gloo", "--same-gpu", "--steps", "5", "--dims", "120", "16", "16", "--transport", "staged"], 29612))


@pytest.mark.skipif("_ngpu() < 2")
def test_two_slabs_nccl():
    _check(_run(2, ["--backend", "nccl", "--steps", "6", "--transport", "nccl"], 29613))


@pytest.mark.skipif("_ngpu() < 2")
def test_two_slabs_p2p_ipc_over_nvlink():
    """One process per GPU, peer mode: CUDA IPC + NVLink peer stores (the bench's multi-GPU path)."""
    res = _run(2, ["--backend", "nccl", "--steps", "6", "--transport", "p2p"], 29620)
    _check(res)
    assert res["transport"] == "p2p"


@pytest.mark.skipif("_ngpu() < 4")
def test_four_slabs_p2p_ipc_over_nvlink_rebalanced():
    res = _run(4, ["--backend", "nccl", "--steps", "120", "--transport", "p2p", "--flow", "--dims", "320", "10", "8", "--rebalance-every", "2"], 29621)
    _check(res)
    assert res["n_rebalances"] >= 3 and max(res["owned"]) / (sum(res["owned"]) / 4) <= 1.08, res["owned"]


@pytest.mark.skipif("_ngpu() < 2")
def test_two_slabs_nccl_overlapped_exchange():
    """Boundary columns first, exchange on a second stream while the interior computes: same bits."""
    _check(_run(2, ["--backend", "nccl", "--steps", "6", "--transport", "nccl"], 29615, env={"PBF_SLAB_OVERLAP": "1"}))


@pytest.mark.skipif("_ngpu() < 4")
def test_four_slabs_nccl():
    _check(_run(4, ["--backend", "nccl", "--steps", "6", "--dims", "160", "20", "20", "--transport", "nccl"], 29614))
