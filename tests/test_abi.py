"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/*.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for hdr in ("pbf_b200.h", "pbf_b200_slab.h", "pbf_b200_multi.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(pbf_[a-z0-9_]+)\s*\(", text))
    return syms


def test_library_exports_every_declared_symbol():
    from fluid_b200 import api
    lib = api.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_library_is_sm100a_cuda_and_independent_of_oracle_and_torch():
    from fluid_b200 import build
    so = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "oracle" not in ldd
    # no product source mentions the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fluid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "pbf_oracle" not in text and "import helpers" not in text, f


def test_default_params_are_the_reference_macros():
    """particles.cpp:24-44 and the hard-coded box of clamp()/clamp_response() (59-83)."""
    from fluid_b200 import api
    p = api.default_params()
    assert (p.h, p.dt, p.eps_relax, p.k_corr, p.dq_ratio, p.visc_c, p.vort_eps, p.gravity_y) == (0.3, 0.016, 2.0, 0.0001, 0.1, 0.001, 0.001, -10.0)
    assert (p.n_corr, p.iterations, p.rest_density) == (4, 12, 1000.0)
    assert list(p.box_min) == [-1.0, 0.0, -1.0] and list(p.box_max) == [1.0, 1.49, 1.0]
    assert (p.y_light, p.z_front, p.xsph_mode, p.enable_vorticity, p.enable_xsph) == (1.49, 1.0, 0, 1, 1)
    # the oracle's mirror of the struct has the same layout
    import helpers
    assert C.sizeof(api.PbfParams) == C.sizeof(helpers.PbfParams)
    q = helpers.default_params()
    assert bytes(p) == bytes(q)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fluid_b200 import api
    lib = api.load_library()
    assert lib.pbf_device_count() == 0
    with pytest.raises(api.PbfError) as e:
        api.Solver(api.default_params())
    assert e.value.code == api.PBF_ERR_NO_DEVICE


def test_grid_dims_and_columns_on_host():
    """pbf_grid_dims / pbf_cell_columns are pure host functions (used by the slab planner)."""
    from fluid_b200 import api, slab
    lib = slab._bind(api.load_library())
    p = api.default_params(box_min=(0, 0, 0), box_max=(120.0, 30.0, 20.1))
    dims = (C.c_int * 3)()
    assert lib.pbf_grid_dims(C.byref(p), C.byref(dims)) == 0
    cell = np.float32(0.3) * (np.float32(1) + np.float32(1 / 256))
    zsub = int(os.environ.get("PBF_ZSUB", "8"))     # thin cells along z (DevParams::zsub)
    assert list(dims) == [int(np.floor(120.0 / float(cell))) + 1, int(np.floor(30.0 / float(cell))) + 1, (int(np.floor(20.1 / float(cell))) + 1) * zsub]
    x = np.array([[0.0, 1, 1], [0.30, 1, 1], [0.302, 1, 1], [119.99, 1, 1], [500.0, 1, 1]])
    col = np.empty(5, dtype=np.int32)
    assert lib.pbf_cell_columns(C.byref(p), 5, x.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p)) == 0
    assert list(col) == [0, 0, 1, int(np.floor(np.float32(119.99) / cell)), dims[0] - 1]


def test_obstacle_hierarchy_build_on_host():
    """pbf_debug_build_bvh (pure host): the hierarchy pbf_set_obstacle_triangles hands to the device.  Every triangle
    sits in exactly one leaf of at most 4, every node's box contains the boxes of its whole subtree, children are
    adjacent, and the depth stays far below the traversal stack (40) — also for 5000 identical triangles."""
    import helpers as H
    from fluid_b200 import api
    lib = api.load_library()
    lib.pbf_debug_build_bvh.restype = C.c_int
    lib.pbf_debug_build_bvh.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    mesh = np.concatenate([H.uv_sphere_mesh((-0.5, 0.32, 0.5), 0.3, 24, 48), H.heightfield_mesh(-1.05, 1.05, -1.05, 1.05, 40, 40)])
    mesh = mesh[rng.permutation(len(mesh))]
    cases = {"mixed": mesh, "one": mesh[:1], "five": mesh[:5], "identical": np.tile(mesh[:1], (5000, 1))}
    for name, tris in cases.items():
        tris = np.ascontiguousarray(tris, dtype=np.float64)
        n = len(tris)
        nodes = np.empty((2 * n + 8, 8), dtype=np.float32); order = np.empty(n, dtype=np.uint32)
        nn = C.c_size_t(); depth = C.c_int()
        assert lib.pbf_debug_build_bvh(n, tris.ctypes.data, nodes.ctypes.data, len(nodes), order.ctypes.data, C.byref(nn), C.byref(depth)) == 0, name
        nodes = nodes[:nn.value]
        ia = nodes[:, 3].copy().view(np.int32); ib = nodes[:, 7].copy().view(np.int32)
        assert sorted(order.tolist()) == list(range(n)), name
        assert depth.value <= 2 + int(np.ceil(np.log2(max(n, 2)))) and depth.value <= 40, (name, depth.value)
        v = tris[:, :9].reshape(n, 3, 3).astype(np.float32)
        tlo, thi = v.min(axis=1), v.max(axis=1)
        seen = np.zeros(n, dtype=int)

        def check(node, lo, hi):                      # returns the box of the subtree, asserts containment on the way
            a, b = int(ia[node]), int(ib[node])
            nlo, nhi = nodes[node, 0:3], nodes[node, 4:7]
            assert np.all(nlo >= lo) and np.all(nhi <= hi), name
            if b > 0:
                assert b <= 4
                ids = order[a:a + b]; seen[ids] += 1
                assert np.all(tlo[ids] >= nlo) and np.all(thi[ids] <= nhi), name
                return
            assert 0 < a and a + 1 < len(nodes)
            check(a, nlo, nhi); check(a + 1, nlo, nhi)
        check(0, np.full(3, -np.inf, dtype=np.float32), np.full(3, np.inf, dtype=np.float32))
        assert np.all(seen == 1), name
    nn = C.c_size_t(); depth = C.c_int()
    assert lib.pbf_debug_build_bvh(len(mesh), np.ascontiguousarray(mesh).ctypes.data, None, 0, None, C.byref(nn), C.byref(depth)) == api.PBF_ERR_CAPACITY
    assert nn.value > len(mesh) // 4


def test_hierarchy_walk_equals_scan_in_a_model_of_the_device_code():
    """Model (numpy, fp64) of ex_mesh_hit / ex_tri_test (fluid_b200/csrc/pbf_device.cuh): the walk over the host-built
    hierarchy, pruning with the SAME margin pbf_set_obstacle_triangles puts on the node boxes, must return the winner
    of the scan over all triangles for every segment — including starts inside the contact skin, a hair behind a
    plane, grazing directions and equally near hits on duplicated triangles (larger index wins).  This pins the
    reasoning behind the margin (every accepted hit lies within tol_ray of the segment); the CUDA code itself is
    checked bit for bit against the oracle's scan on the GPU (tests/test_gpu_parity.py)."""
    import helpers as H
    from fluid_b200 import api
    lib = api.load_library()
    lib.pbf_debug_build_bvh.restype = C.c_int
    lib.pbf_debug_build_bvh.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11)
    base = np.concatenate([H.uv_sphere_mesh((0.2, 0.5, -0.1), 0.4, 12, 24), H.heightfield_mesh(-1.0, 1.0, -1.0, 1.0, 16, 16)])
    dup = base[rng.choice(len(base), 200, replace=False)]
    tris = np.ascontiguousarray(np.concatenate([base, dup])[rng.permutation(len(base) + 200)], dtype=np.float64)
    n = len(tris)
    nodes = np.empty((2 * n + 8, 8), dtype=np.float32); order = np.empty(n, dtype=np.uint32)
    nn = C.c_size_t(); depth = C.c_int()
    assert lib.pbf_debug_build_bvh(n, tris.ctypes.data, nodes.ctypes.data, len(nodes), order.ctypes.data, C.byref(nn), C.byref(depth)) == 0
    ia = nodes[:nn.value, 3].copy().view(np.int32); ib = nodes[:nn.value, 7].copy().view(np.int32)
    nodes = nodes[:nn.value].astype(np.float64)
    # the constants of fill_dev_params / pbf_set_obstacle_triangles for the default box (largest coordinate 1.49)
    h = 0.3; ulp = 1.49 * 1.1920928955078125e-07
    skin = max(1e-5 * h, 16 * ulp); tol_n = max(1e-4 * h, 64 * ulp) + skin; tol_ray = max(0.05 * h, 16 * max(1e-4 * h, 64 * ulp))
    margin = 1e-2 * h + tol_ray + 2 * skin
    nlo = nodes[:, 0:3] - margin - 1e-5 * np.abs(nodes[:, 0:3]); nhi = nodes[:, 4:7] + margin + 1e-5 * np.abs(nodes[:, 4:7])
    p1 = tris[:, 0:3]; e1 = tris[:, 3:6] - p1; e2 = tris[:, 6:9] - p1
    ng = np.cross(e1, e2); ngl = np.linalg.norm(ng, axis=1)
    sg = np.where(np.einsum("ij,ij->i", ng, tris[:, 9:12] + tris[:, 12:15] + tris[:, 15:18]) < 0, -1.0, 1.0)
    BT, TOL_T = 1e-6, 1e-4 * h

    def tri_test(k, o, d, max_t, best):
        s = o - p1[k]; s1 = np.cross(d, e2[k]); s2 = np.cross(s, e1[k]); dd = s1 @ e1[k]
        if not (sg[k] * dd > 0):
            return None
        t = (s2 @ e2[k]) / dd
        t -= min(skin * ngl[k] / abs(dd), tol_ray)
        if t < 0:
            if t >= -TOL_T or (t >= -tol_ray and -t * abs(dd) <= tol_n * ngl[k]):
                t = 0.0
            else:
                return None
        if t > max_t or (t == max_t and best >= 0 and k < best):
            return None
        u = (s1 @ s) / dd; v = (s2 @ d) / dd; w = 1 - u - v
        if min(u, v, w) < -BT or max(u, v, w) > 1 + BT:
            return None
        return t

    def scan(o, d, max_t):
        best = -1
        for k in range(n):
            t = tri_test(k, o, d, max_t, best)
            if t is not None:
                max_t, best = t, k
        return best, max_t

    def walk(o, d, max_t):
        best = -1; stack = [0]; visited = 0
        while stack:
            nd = stack.pop()
            q = o + max_t * d
            lo, hi = np.minimum(o, q), np.maximum(o, q)
            if np.any(hi < nlo[nd]) or np.any(lo > nhi[nd]):
                continue
            if ib[nd] > 0:
                for slot in range(ia[nd], ia[nd] + ib[nd]):
                    k = int(order[slot]); visited += 1
                    t = tri_test(k, o, d, max_t, best)
                    if t is not None:
                        max_t, best = t, k
            else:
                stack += [ia[nd] + 1, ia[nd]]
        return best, max_t, visited
    hits = contacts = 0; visited_total = 0; trials = 400
    cen = (p1 + tris[:, 3:6] + tris[:, 6:9]) / 3; nh = (ng / ngl[:, None]) * sg[:, None]
    for trial in range(trials):
        k = int(rng.integers(n))
        kind = trial % 4
        off = [0.05, 0.5 * skin, -0.5 * tol_n, 0.02][kind]                    # far in front / inside the skin / a hair behind / near
        noise = 0.01 * rng.normal(size=3) if kind != 3 else np.zeros(3)
        if kind in (1, 2):
            noise -= (noise @ nh[k]) * nh[k]                                    # stay at that distance from the plane
        o = cen[k] + off * nh[k] + noise
        d = -nh[k] + ([0.3, 0.3, 0.3, 30.0][kind]) * rng.normal(size=3)       # kind 3: grazing
        d /= np.linalg.norm(d)
        max_t = float(rng.uniform(0.001, 0.12))
        b1, t1 = scan(o, d, max_t); b2, t2, vis = walk(o, d, max_t)
        assert (b1, t1) == (b2, t2), (trial, kind, b1, t1, b2, t2)
        hits += b1 >= 0; contacts += (b1 >= 0 and t1 == 0.0); visited_total += vis
    assert hits > trials // 4 and contacts > 20
    assert visited_total < 0.1 * trials * n                                   # and the walk does prune


def test_slab_planners_on_host():
    """pbf_partition_columns / pbf_plan_rebalance (pure host, fluid_b200/csrc/pbf_multi.cpp): the planner of pbf_create_multi
    equals slab.partition_columns, and a re-balancing move keeps every boundary strictly between its old neighbours (one
    migration hop), moves towards the equal-count partition and never carries more than the message budget."""
    from fluid_b200 import api, slab
    lib = api.load_library()
    lib.pbf_partition_columns.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.pbf_plan_rebalance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_double, C.c_void_p, C.POINTER(C.c_double)]
    rng = np.random.default_rng(3)
    for trial in range(200):
        ncol = int(rng.integers(1, 60)); world = int(rng.integers(1, min(ncol, 8) + 1))
        hist = rng.integers(0, 1000, size=ncol).astype(np.uint64)
        if trial % 5 == 0:
            hist[rng.integers(0, ncol):] = 0             # fluid only in a part of the tank
        out = np.zeros(world + 1, dtype=np.int32)
        assert lib.pbf_partition_columns(hist.ctypes.data, ncol, world, out.ctypes.data) == 0
        assert list(out) == slab.partition_columns(hist, world)
    # a dam break: all the fluid in the left third at first, then spread evenly; the slabs follow step by step
    ncol, world = 90, 4
    hist0 = np.zeros(ncol, dtype=np.uint64); hist0[:30] = 1000
    b = np.zeros(world + 1, dtype=np.int32); lib.pbf_partition_columns(hist0.ctypes.data, ncol, world, b.ctypes.data)
    assert list(b) == [0, 8, 15, 22, 90] or b[-1] == 90
    hist1 = np.full(ncol, 333, dtype=np.uint64)
    imb = C.c_double()
    for it in range(200):
        owned_hist = hist1.copy()
        nb = np.zeros(world + 1, dtype=np.int32)
        ch = lib.pbf_plan_rebalance(owned_hist.ctypes.data, ncol, world, b.ctypes.data, 1500, 1.05, nb.ctypes.data, C.byref(imb))
        assert ch in (0, 1)
        assert nb[0] == 0 and nb[-1] == ncol and np.all(np.diff(nb) >= 1)
        for k in range(1, world):
            assert b[k - 1] < nb[k] < b[k + 1]                                     # one hop
            lo, hi = sorted((int(b[k]), int(nb[k])))
            assert owned_hist[lo:hi].sum() <= 1500                                  # fits the message budget
        if ch == 0:
            break
        b = nb
    owned = np.array([hist1[b[r]:b[r + 1]].sum() for r in range(world)], dtype=np.float64)
    assert owned.max() / owned.mean() <= 1.05 and it < 60, (list(b), it)
    # balanced input: nothing moves
    nb = np.zeros(world + 1, dtype=np.int32)
    assert lib.pbf_plan_rebalance(hist1.ctypes.data, ncol, world, b.ctypes.data, 1500, 1.05, nb.ctypes.data, C.byref(imb)) == 0 and list(nb) == list(b)
