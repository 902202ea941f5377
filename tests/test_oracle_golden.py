"""Pin the CPU oracle: Oracle<double>(reference-order XSPH, triangle walls) must reproduce the
UNMODIFIED reference (particles.cpp:250-297) bit-for-bit — full state and ordered neighbour lists —
against the committed fixtures (tests/golden/, made by make_golden.py from oracle/_ref/ref_harness)
and, when the compiled reference is present, against a live run."""
import os
import subprocess

import numpy as np
import pytest

from helpers import (COLLIDE_BOX, COLLIDE_TRIANGLES, GOLDEN, SEARCH_BRUTE, SEARCH_GRID, XSPH_JACOBI,
                     XSPH_REFERENCE, Oracle, default_params, have_reference_binary, read_dump,
                     ref_harness_path, state_sha, write_bin_scene)

SHIPPED = ["p", "spheres_p"]
JITTER = ["two_blocks", "sparse", "corner", "front"]
SPHERES = ["sphere_drop", "sphere_hit"]     # the CBspheres obstacle spheres in the reference's BVH
MESHES = ["mesh_drop"]                      # obstacle triangles (a cuboid and a wedge) in the reference's BVH


def _load(name):
    if name in SHIPPED:
        sc = np.load(os.path.join(GOLDEN, f"scene_{name}.npz"))
        ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
        return sc["pos"], sc["vel"], float(sc["rho0"]), ref
    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz" if name in SPHERES + MESHES else f"ref_jitter_{name}.npz"))
    return ref["pos"], ref["vel"], float(ref["rho0"]), ref


def _oracle(rho0, ref, cmode, search=SEARCH_GRID, xsph=XSPH_REFERENCE):
    o = Oracle(default_params(rest_density=rho0, xsph_mode=xsph), 64, cmode, search)
    if "spheres" in ref.files:
        o.set_spheres(ref["spheres"])
    if "tris" in ref.files:
        o.set_triangles(ref["tris"])
    return o


def test_constants():
    # the reference's literal H2 0.09 equals H*H in both working precisions (particles.cpp:26-28)
    assert 0.3 * 0.3 == 0.09
    assert np.float32(0.3) * np.float32(0.3) == np.float32(0.09)
    # tensile scale 1/poly6(0,0,0.1H) (particles.cpp:151), value recorded in SURVEY.md §8a
    h = 0.3; r2 = (0.1 * h) ** 2; t = h * h - r2
    w = 1.56668147106 * (t * t * t) / h ** 9
    assert abs(1.0 / w - 0.017761411379071678) < 1e-15


@pytest.mark.parametrize("name", SHIPPED + JITTER + SPHERES + MESHES)
@pytest.mark.parametrize("search", [SEARCH_BRUTE, SEARCH_GRID])
def test_oracle_matches_reference_fixture(name, search):
    """With obstacle spheres the oracle walks a restatement of the reference's own BVH (bvh.cpp:48-192), because
    clamp() takes the first primitive the traversal finds, not the nearest (particles.cpp:76)."""
    pos, vel, rho0, ref = _load(name)
    steps = len(ref["sha"])
    o = _oracle(rho0, ref, COLLIDE_TRIANGLES, search)
    o.upload(pos, vel); o.estimate_densities()
    keep = set(int(k) for k in ref["keep"])
    for s in range(steps):
        o.step()
        P, V, R = o.download()
        assert state_sha(P, V, R) == str(ref["sha"][s]), f"{name}: state differs from reference at step {s}"
        row, col = o.neighbors()
        assert len(col) == int(ref["pairs"][s])
        if s == 0:
            assert np.array_equal(np.diff(row.astype(np.int64)), ref["nbr_counts_0"])
            assert np.array_equal(col.astype(np.int32), ref["nbr_col_0"])   # ordered lists, exact
        if s in keep:
            st = ref[f"state_{s}"]
            assert np.array_equal(P, st[:, 0:3]) and np.array_equal(V, st[:, 3:6]) and np.array_equal(R, st[:, 6])
        a, b, _ = o.stats()
        ta, tb = ref["avg_rho_text"][s]
        assert f"{a:.6g}" == ta and f"{b:.6g}" == tb    # the two numbers the reference prints per step


def test_survey_golden_values():
    """SURVEY.md §8c table (values captured from the unmodified reference during the survey)."""
    ref = np.load(os.path.join(GOLDEN, "ref_p.npz"))
    assert [str(ref["sha"][s]) for s in (0, 1, 19)] == ["dc17aefca6e0d81e", "91a4408ef8e9c1ef", "57e229f96ef1800b"]
    assert [int(ref["pairs"][s]) for s in (0, 1, 19)] == [15988, 14012, 13548]
    assert tuple(ref["avg_rho_text"][0]) == ("139.82", "145.755")
    st0 = ref["state_0"][0]
    assert st0[0] == -0.8527234379438056 and st0[1] == 0.528251353076964 and st0[6] == 137.5904451565677
    ref = np.load(os.path.join(GOLDEN, "ref_spheres_p.npz"))
    assert [str(ref["sha"][s]) for s in (0, 1, 19)] == ["a6828113d0842938", "ac3fc0fd5f86a8ee", "3b449a32a53f34dd"]
    assert [int(ref["pairs"][s]) for s in (0, 1, 19)] == [157750, 125400, 153134]


@pytest.mark.parametrize("name", SHIPPED + JITTER + SPHERES)
def test_analytic_box_equals_triangles_fp64(name):
    """Chain of trust, second arrow: the analytic box with the fp32 contact rules (one-sided planes,
    exact axis normals, sticky virtual planes) is the same operator as the
    reference's triangle walls in fp64.  Teacher-forced per step from the reference's own state:
    neighbour lists identical, state equal up to the rounding of the hit distance t
    (Moller-Trumbore vs (plane-o)/d: a few ulp, amplified by at most the 12 iterations).
    Obstacle spheres: the one-sided sphere rule (blocks only motion into the sphere, entry root clamped to
    t >= 0, nearest hit) against the reference's Sphere::test + BVH any-hit order."""
    pos, vel, rho0, ref = _load(name)
    keep = sorted(int(k) for k in ref["keep"])
    checked = 0
    for k, s in enumerate(keep[:-1]):
        if keep[k + 1] != s + 1:
            continue
        st = ref[f"state_{s}"]; nxt = ref[f"state_{s + 1}"]
        res = []
        for cm in (COLLIDE_TRIANGLES, COLLIDE_BOX):
            o = _oracle(rho0, ref, cm)
            o.upload(st[:, 0:3], st[:, 3:6]); o.step()
            res.append(o.download() + (o.neighbors(),))
        assert np.array_equal(res[0][0], nxt[:, 0:3])              # triangles == reference (teacher-forced)
        assert np.array_equal(res[0][3][1], res[1][3][1])           # identical ordered neighbour lists
        assert np.abs(res[0][0] - res[1][0]).max() < 1e-11
        assert np.abs(res[0][2] - res[1][2]).max() < 1e-8 * rho0
        checked += 1
    assert checked >= 1
    if name in SPHERES:     # the fixture really exercises sphere contacts: particles within 1e-6 of a sphere surface
        P = ref[f"state_{keep[-1]}"][:, 0:3]
        near = sum(int((np.abs(np.linalg.norm(P - c[:3], axis=1) - c[3]) < 1e-6).sum()) for c in ref["spheres"])
        assert near >= 10, near


def test_one_sided_triangles_vs_reference_triangles():
    """Obstacle triangles.  The reference's triangles are two-sided and its clamp() takes the first primitive the BVH
    traversal finds, so it lets particles INTO the obstacles (9 in the cuboid, 74 under the wedge after 80 steps) and
    traps particles that end up 1e-11 behind a face.  The one-sided rule (GPU) fixes that, so the two can only be
    compared where the reference does not misbehave:
      * before the first anomaly (steps 0-12 of the fixture) whole steps agree to 1e-11, teacher-forced;
      * afterwards the collision operator alone (iterations = 0) still agrees for all but a handful of particles
        in 1e-11-contact with a face;
      * free-running, the one-sided rule keeps every particle outside both obstacles, in fp64 and in fp32."""
    pos, vel, rho0, ref = _load("mesh_drop")
    keep = sorted(int(k) for k in ref["keep"])
    for k, s in enumerate(keep[:-1]):
        if keep[k + 1] != s + 1:
            continue
        st = ref[f"state_{s}"]; nxt = ref[f"state_{s + 1}"]
        if s + 1 <= 12:
            o = _oracle(rho0, ref, COLLIDE_BOX); o.upload(st[:, 0:3], st[:, 3:6]); o.step()
            assert np.abs(o.download()[0] - nxt[:, 0:3]).max() < 1e-11, s
        xp = []
        for cm in (COLLIDE_TRIANGLES, COLLIDE_BOX):
            prm = default_params(rest_density=rho0, xsph_mode=XSPH_REFERENCE, iterations=0)
            o = Oracle(prm, 64, cm, SEARCH_GRID); o.set_triangles(ref["tris"])
            o.upload(st[:, 0:3], st[:, 3:6]); o.step()
            from helpers import ARRAY_XPRED
            xp.append(o.array(ARRAY_XPRED))
        differ = int((np.linalg.norm(xp[0] - xp[1], axis=1) > 1e-9).sum())
        assert differ <= 10, (s, differ)

    def inside(P, tol=1e-5):
        box = (P[:, 0] > -0.7 + tol) & (P[:, 0] < -0.2 - tol) & (P[:, 1] < 0.4 - tol) & (P[:, 2] > -0.7 + tol) & (P[:, 2] < -0.2 - tol)
        wedge = ((P[:, 0] > 0.15 + tol) & (P[:, 0] < 0.9 - tol) & (P[:, 2] > 0.2 + tol) & (P[:, 2] < 0.8 - tol) &
                 (P[:, 1] < (P[:, 0] - 0.15) / 0.75 * 0.5 - tol))
        return int(box.sum()), int(wedge.sum())
    assert sum(inside(ref["state_39"][:, 0:3])) >= 20          # the reference itself leaks
    for prec in (64, 32):
        o = Oracle(default_params(rest_density=rho0, xsph_mode=XSPH_JACOBI), prec, COLLIDE_BOX, SEARCH_GRID)
        o.set_triangles(ref["tris"]); o.upload(pos, vel)
        for _ in range(4):
            o.step(15)
            assert inside(o.download()[0]) == (0, 0), prec


@pytest.mark.parametrize("name", SHIPPED)
def test_jacobi_xsph_changes_only_velocity(name):
    """Quirk Q11: Jacobi XSPH (what the GPU computes) gives the same positions and densities as the
    reference's in-index-order XSPH within a step; only velocities differ."""
    pos, vel, rho0, ref = _load(name)
    out = []
    for mode in (XSPH_REFERENCE, XSPH_JACOBI):
        o = Oracle(default_params(rest_density=rho0, xsph_mode=mode), 64, COLLIDE_TRIANGLES, SEARCH_GRID)
        o.upload(pos, vel); o.step(); out.append(o.download())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2])
    assert not np.array_equal(out[0][1], out[1][1])


@pytest.mark.skipif(not have_reference_binary(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference(tmp_path):
    """Live run of the unmodified reference on a fresh jittered input not in the fixtures."""
    rng = np.random.default_rng(2024)
    from helpers import lattice_block
    pos, vel = lattice_block(7, 9, 6, origin=(-0.6, 0.05, -0.95), spacing=0.1, v0=(0.4, -2.0, -0.7), jitter=0.002, seed=5)
    vel = vel + rng.normal(0, 0.05, size=vel.shape)
    scene = str(tmp_path / "s.bin"); dump = str(tmp_path / "d.bin")
    write_bin_scene(scene, pos, vel, 700.0)
    subprocess.run([ref_harness_path(), "--bin", scene, "--steps", "10", "--out", dump, "--quiet"], check=True)
    d = read_dump(dump)
    o = Oracle(default_params(rest_density=700.0, xsph_mode=XSPH_REFERENCE), 64, COLLIDE_TRIANGLES, SEARCH_GRID)
    o.upload(pos, vel)
    for s in range(10):
        o.step()
        P, V, R = o.download()
        assert np.array_equal(P, d[s]["state"][:, 0:3]) and np.array_equal(V, d[s]["state"][:, 3:6])
        assert np.array_equal(R, d[s]["state"][:, 6])
        assert np.array_equal(o.neighbors()[1].astype(np.int32), d[s]["col"])


@pytest.mark.parametrize("name", ["two_blocks", "sparse"])
def test_density_field_matches_reference(name):
    """Particles::estimateDensityAt (particles.cpp:446-453), the marching-cubes field: bit-exact."""
    ref = np.load(os.path.join(GOLDEN, f"ref_density_{name}.npz"))
    o = Oracle(default_params(rest_density=float(ref["rho0"])), 64, COLLIDE_TRIANGLES, SEARCH_GRID)
    o.upload(ref["pos"], ref["vel"])
    assert np.array_equal(o.density_at(ref["q"]), ref["density"])


# ---- marching-cubes surface ----------------------------------------------------

def test_polygonise_table_matches_reference_for_all_256_patterns():
    """The marching-cubes triangle table (fluid_b200/csrc/pbf_mc_table.h, used by the oracle and by the CUDA surfacer)
    against the reference's own polygonise() (marching.cpp:17-380) on the unit cell for every sign pattern: same
    triangles, same order, same vertex order."""
    from helpers import oracle_polygonise_case
    fx = np.load(os.path.join(GOLDEN, "mc_cases.npz"))
    counts, verts = fx["counts"], fx["verts"]
    assert counts.sum() == 820 and counts.max() == 5 and counts[0] == 0 and counts[255] == 0
    off = 0
    for c in range(256):
        got = oracle_polygonise_case(c)
        want = verts[off:off + counts[c]]; off += counts[c]
        assert got.shape == want.shape and np.array_equal(got, want), c


@pytest.mark.parametrize("name", ["p", "spheres_p"])
def test_surface_matches_reference_fixture(name):
    """Oracle restatement of Particles::getSurfacePrims (lattice, polygonise, vertexInterp, getVertexNormal;
    particles.cpp:309-418, marching.cpp) == the unmodified reference's triangle soup on the reference's own state,
    bit for bit: vertices and normals, in order."""
    fx = np.load(os.path.join(GOLDEN, f"ref_surface_{name}.npz"))
    st, want, rho0 = fx["state"], fx["tris"], float(fx["rho0"])
    o = Oracle(default_params(rest_density=rho0), 64, COLLIDE_TRIANGLES, SEARCH_GRID)
    o.upload(st[:, 0:3], st[:, 3:6])
    got = o.surface(rho0)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want), f"{int(np.any(got != want, axis=1).sum())} triangles differ, max {np.abs(got - want).max():.3e}"


@pytest.mark.parametrize("offset", [0.0, 100.0])
def test_fp32_triangle_contact_rules_do_not_leak(offset):
    """The fp32-only contact rules of the one-sided triangles (rest a skin in front of the plane, hold what is a hair
    behind it; oracle/pbf_oracle.hpp mesh_hit_onesided) exist because fp32 cannot resolve the reference's 1e-11
    stand-off: without them 637 of 41k particles sat inside a tessellated sphere after 5 steps at |x| ~ 100, and a few
    even at the origin.  With them nothing is inside the polyhedron, at the origin and at |x| ~ 100 (one ulp = 8e-6)
    alike, and fp64 (which never leaked) agrees.  The GPU reproduces the fp32 oracle's x* bit for bit
    (tests/test_gpu_parity.py), so this is also its guarantee."""
    from helpers import lattice_block, uv_sphere_mesh
    pos, vel = lattice_block(30, 24, 30, origin=(offset + 0.1, 0.1, offset + 0.1), jitter=0.001)
    c = np.array([offset + 1.5, 0.0, offset + 1.5]); r = 1.0
    keep = np.linalg.norm(pos - c, axis=1) > r + 0.02
    pos, vel = pos[keep], vel[keep]
    tris = uv_sphere_mesh(c, r, 16, 32)
    e1 = tris[:, 3:6] - tris[:, 0:3]; e2 = tris[:, 6:9] - tris[:, 0:3]
    ng = np.cross(e1, e2); ng /= np.linalg.norm(ng, axis=1)[:, None]

    def inside(P):                                   # particles strictly inside the convex polyhedron, by more than 1e-4
        near = np.flatnonzero(np.linalg.norm(P - c, axis=1) < r * 1.0005)
        if len(near) == 0:
            return 0
        sd = np.einsum("pfk,fk->pf", P[near][:, None, :] - tris[None, :, 0:3], ng).max(axis=1)
        return int((sd < -1e-4).sum())
    prm = default_params(rest_density=700.0, xsph_mode=XSPH_JACOBI, box_min=(offset, 0, offset), box_max=(offset + 9, 12, offset + 3.1),
                         y_light=12.0, z_front=offset + 3.1)
    for prec in (32, 64):
        o = Oracle(prm, prec, COLLIDE_BOX, SEARCH_GRID); o.set_triangles(tris); o.upload(pos, vel)
        for _ in range(4):
            o.step(1)
            assert inside(o.download()[0]) == 0, (offset, prec)
    if offset > 0:                                   # and without the rules the fp32 oracle does leak there
        os.environ["PBF_ORACLE_NO_FP32_CONTACT_RULES"] = "1"
        try:
            o = Oracle(prm, 32, COLLIDE_BOX, SEARCH_GRID); o.set_triangles(tris); o.upload(pos, vel)
        finally:
            del os.environ["PBF_ORACLE_NO_FP32_CONTACT_RULES"]
        o.step(4)
        assert inside(o.download()[0]) > 0
