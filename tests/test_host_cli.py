"""The C++ host adapter (fluid_b200/host): XML import like Application::load_particles
(application.cpp:302-344) and the windowless -p/-d loop (pathtracer.cpp:444-480)."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, read_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fluid_b200", "host")


def _build():
    from fluid_b200 import build
    build.build()
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return os.path.join(HOST, "pbf_run")


def _write_xml(path, pos, vel, rho0):
    with open(path, "w") as f:
        f.write("<?xml version=\"1.0\"?>\n<!-- generated -->\n<particles>\n  <density>%r</density>\n  <ps>\n" % float(rho0))
        for p, v in zip(pos, vel):
            f.write("    <particle>\n      <pos>%r %r %r</pos>\n      <v>%r %r %r</v>\n    </particle>\n" % (*map(float, p), *map(float, v)))
        f.write("  </ps>\n</particles>\n")


def test_xml_loader_matches_reference_parse(tmp_path):
    exe = _build()
    sc = np.load(os.path.join(GOLDEN, "scene_spheres_p.npz"))
    xml = str(tmp_path / "s.xml")
    _write_xml(xml, sc["pos"], sc["vel"], float(sc["rho0"]))
    out = json.loads(subprocess.run([exe, "-p", xml, "--parse-only"], check=True, capture_output=True, text=True).stdout)
    assert out["n"] == 2106 and out["rho0"] == 700.0
    assert out["sum_pos"] == float(np.sum(sc["pos"].reshape(-1).cumsum()[-1:])) or abs(out["sum_pos"] - sc["pos"].sum()) < 1e-9
    assert out["sum_vel"] == -2106.0
    # malformed files are reported, not crashed on
    bad = str(tmp_path / "bad.xml"); open(bad, "w").write("<notparticles/>")
    r = subprocess.run([exe, "-p", bad, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "XML error" in r.stdout


@pytest.mark.gpu
def test_cli_run_equals_python_api_and_tracks_reference(tmp_path):
    exe = _build()
    ref = np.load(os.path.join(GOLDEN, "ref_jitter_two_blocks.npz"))
    xml = str(tmp_path / "s.xml"); dump = str(tmp_path / "d.bin")
    _write_xml(xml, ref["pos"], ref["vel"], float(ref["rho0"]))
    r = subprocess.run([exe, "-p", xml, "--steps", "2", "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("avg rho:") == 2
    d = read_dump(dump)
    from fluid_b200 import api
    g = api.Solver(api.default_params(rest_density=float(ref["rho0"])))
    g.upload(ref["pos"], ref["vel"]); g.estimate_densities(); g.step(1)
    P, V, R = g.download()
    assert np.array_equal(d[0]["state"][:, 0:3], P) and np.array_equal(d[0]["state"][:, 3:6], V) and np.array_equal(d[0]["state"][:, 6], R)
    # step 0 against the unmodified reference's output
    dx = np.linalg.norm(d[0]["state"][:, 0:3] - ref["state_0"][:, 0:3], axis=1)
    assert np.percentile(dx, 50) <= 1e-5 and np.percentile(dx, 99) <= 5e-3 and dx.max() <= 1e-1
    # obstacle spheres through the host adapter (Particles::setObstacleSpheres) == the Python binding
    sref = np.load(os.path.join(GOLDEN, "ref_sphere_hit.npz"))
    _write_xml(xml, sref["pos"], sref["vel"], float(sref["rho0"]))
    cmd = [exe, "-p", xml, "--steps", "3", "--dump", dump, "--quiet"]
    for c in sref["spheres"]:
        cmd += ["--sphere"] + [repr(float(v)) for v in c]
    assert subprocess.run(cmd, capture_output=True, text=True).returncode == 0
    d = read_dump(dump)
    g = api.Solver(api.default_params(rest_density=float(sref["rho0"])))
    g.set_obstacle_spheres(sref["spheres"]); g.upload(sref["pos"], sref["vel"]); g.estimate_densities(); g.step(3)
    P, V, R = g.download()
    assert np.array_equal(d[2]["state"][:, 0:3], P) and np.array_equal(d[2]["state"][:, 3:6], V)
    # restart files: 2 steps + checkpoint + 2 more steps == 4 uninterrupted steps, bit for bit (spheres travel along)
    ck = str(tmp_path / "state.ckpt"); dump2 = str(tmp_path / "d2.bin")
    base = [exe, "--quiet"]
    sph = cmd[cmd.index("--sphere"):]
    assert subprocess.run(base + ["-p", xml, "--steps", "4", "--dump", dump] + sph, capture_output=True).returncode == 0
    assert subprocess.run(base + ["-p", xml, "--steps", "2", "--save-state", ck] + sph, capture_output=True).returncode == 0
    assert open(ck, "rb").read(8) == b"PBFCKPT2"
    r = subprocess.run(base + ["--load-state", ck, "--steps", "2", "--dump", dump2], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d4, d22 = read_dump(dump), read_dump(dump2)
    assert np.array_equal(d4[3]["state"][:, :7], d22[1]["state"][:, :7])
    open(ck, "r+b").write(b"NOTACKPT")
    r = subprocess.run(base + ["--load-state", ck, "--steps", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "checkpoint error" in r.stdout
    # obstacle triangles travel in the restart file too: 2 + 2 steps without --tris on the second leg == 4 steps
    import helpers as H
    keep = sref["pos"][:, 1] >= 0.6
    _write_xml(xml, sref["pos"][keep], sref["vel"][keep], float(sref["rho0"]))
    trif = str(tmp_path / "mesh.tris")
    H.write_tris(trif, H.uv_sphere_mesh((-0.5, 0.3, 0.5), 0.25, 10, 20))
    assert subprocess.run(base + ["-p", xml, "--steps", "4", "--dump", dump, "--tris", trif], capture_output=True).returncode == 0
    assert subprocess.run(base + ["-p", xml, "--steps", "2", "--save-state", ck, "--tris", trif], capture_output=True).returncode == 0
    r = subprocess.run(base + ["--load-state", ck, "--steps", "2", "--dump", dump2], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d4, d22 = read_dump(dump), read_dump(dump2)
    assert np.array_equal(d4[3]["state"][:, :7], d22[1]["state"][:, :7])
    r = subprocess.run(base + ["--load-state", ck, "--steps", "1", "--iterations", "4"], capture_output=True, text=True)
    assert r.returncode == 2 and "--iterations" in r.stdout
    r = subprocess.run([exe, "--load-state", ck, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 2
    _write_xml(xml, ref["pos"], ref["vel"], float(ref["rho0"]))
    # -d 0.05 => ceil(0.05/0.016) = 4 steps (while simulate_time < T, Q18)
    r = subprocess.run([exe, "-p", xml, "-d", "0.05", "--quiet"], capture_output=True, text=True)
    assert json.loads(r.stderr.strip().splitlines()[-1])["steps"] == 4


@pytest.mark.gpu
def test_cli_surface_equals_python_api(tmp_path):
    """Particles::updateSurface of the host adapter (pbf_run --surface) == api.Solver.extract_surface on the same run,
    and within the fixture's tolerance of the unmodified reference's surface of the same scene and step count."""
    exe = _build()
    fx = np.load(os.path.join(GOLDEN, "ref_surface_p.npz"))
    sc = np.load(os.path.join(GOLDEN, "scene_p.npz"))
    rho0, steps = float(sc["rho0"]), 3
    xml = str(tmp_path / "p.xml"); sf = str(tmp_path / "surf.bin")
    _write_xml(xml, sc["pos"], sc["vel"], rho0)
    r = subprocess.run([exe, "-p", xml, "--steps", str(steps), "--surface", sf, "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(sf, dtype=np.uint8)
    nt = int(raw[:8].view(np.int64)[0])
    tris = raw[8:].view(np.float64).reshape(nt, 18)
    assert f"including {nt} marching cube surfacing triangles" in r.stderr
    from fluid_b200 import api
    g = api.Solver(api.default_params(rest_density=rho0))
    g.upload(sc["pos"], sc["vel"]); g.estimate_densities(); g.step(steps)
    want = g.extract_surface(rho0)
    assert tris.shape == want.shape and np.array_equal(tris, want)
    assert nt > 500 and abs(nt - len(fx["tris"])) < 0.5 * len(fx["tris"])     # same scene, 3 vs 12 steps: same order of magnitude
    nrm = np.linalg.norm(tris[:, 9:12], axis=1)
    assert np.all((np.abs(nrm - 1) < 1e-9) | (nrm == 0))


@pytest.mark.gpu
def test_cli_neighbor_count_warnings(tmp_path):
    """Particle::initializeWithNewNeighbors (particles.cpp:165-173) warns on cerr about every particle with fewer than
    18 neighbours; the adapter prints the same lines (first 128 in index order, then a summary) from the counts the
    neighbour build leaves on the device.  p.xml (spacing 0.15): the reference prints 100 such lines per step."""
    exe = _build()
    sc = np.load(os.path.join(GOLDEN, "scene_p.npz")); ref = np.load(os.path.join(GOLDEN, "ref_p.npz"))
    xml = str(tmp_path / "p.xml")
    _write_xml(xml, sc["pos"], sc["vel"], float(sc["rho0"]))
    r = subprocess.run([exe, "-p", xml, "--steps", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stderr.splitlines() if " only has " in l and l.startswith("P(p(")]
    from fluid_b200 import api
    g = api.Solver(api.default_params(rest_density=float(sc["rho0"])))
    g.upload(sc["pos"], sc["vel"]); g.estimate_densities(); g.capture(True); g.step(1)
    _, cnt = g.neighbor_digest()
    low = np.flatnonzero(cnt < 18)
    more = [l for l in r.stderr.splitlines() if "more particles with fewer than 18" in l]
    shown = len(lines)
    assert shown == len(low) if len(low) <= 128 else (0 < shown <= 128 and int(more[0].split("and ")[1].split()[0]) == len(low) - shown)
    assert len(low) == int(np.count_nonzero(ref["nbr_counts_0"] < 18))        # the unmodified reference's own count for this step
    xp = g.array(3)                                                            # PBF_ARRAY_XPRED
    for l, i in zip(lines, low[:shown]):
        assert l.endswith(f" only has {cnt[i]} neighbors.")
        x = [float(t) for t in l[len("P(p("):l.index("),v(")].split(",")]
        assert np.allclose(x, xp[i], rtol=1e-5, atol=1e-6)                     # 6 significant digits on the stream
    # the C ABI itself: a deterministic prefix (by id) of the records, the total beyond the cap still counted
    import ctypes as C
    lib = api.load_library()
    lib.pbf_set_neighbor_alert.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    lib.pbf_get_neighbor_alerts.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    assert lib.pbf_set_neighbor_alert(g.h, 18, 16) == 0
    g.upload(sc["pos"], sc["vel"]); g.step(1)
    ids = np.zeros(16, dtype=np.uint32); cn = np.zeros(16, dtype=np.uint32); tot = C.c_size_t(); wr = C.c_size_t()
    assert lib.pbf_get_neighbor_alerts(g.h, 16, ids.ctypes.data_as(C.c_void_p), cn.ctypes.data_as(C.c_void_p), None, None, C.byref(wr), C.byref(tot)) == 0
    _, cnt2 = g.neighbor_digest()
    low2 = np.flatnonzero(cnt2 < 18)
    m = wr.value
    assert tot.value == len(low2) and 0 < m <= 16 and np.array_equal(ids[:m], low2[:m]) and np.array_equal(cn[:m], cnt2[ids[:m]])


@pytest.mark.gpu
def test_cli_several_gpus_equal_one(tmp_path):
    """`pbf_run --devices a,b,c` (pbfhost::Particles::setDevices -> pbf_create_multi): the reference's load / estimateDensities /
    timeStep loop on x-slabs behind the same object equals the one-device run bit for bit (positions, velocities, densities
    of every step, the load-time densities included), obstacle sphere on a slab boundary, marching-cubes surface too."""
    import torch
    exe = _build()
    from helpers import lattice_block
    ng = torch.cuda.device_count()
    pos, vel = lattice_block(60, 14, 12, origin=(0.1, 0.1, 0.1), spacing=0.1, v0=(0.8, -1.0, 0.0), jitter=0.001, seed=3)
    xml = str(tmp_path / "tank.xml")
    _write_xml(xml, pos, vel, 700.0)
    box = ["--box", "0", "0", "0", "9.3", "3.0", "1.5"]
    sph = ["--sphere", "3.05", "0.5", "0.7", "0.45"]
    outs = {}
    for tag, dev in (("one", None), ("three", ",".join(str(d % ng) for d in range(3)))):
        dump = str(tmp_path / f"{tag}.bin"); surf = str(tmp_path / f"{tag}.surf")
        cmd = [exe, "-p", xml, "--steps", "5", "--dump", dump, "--surface", surf, "--quiet"] + box + sph + (["--devices", dev] if dev else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        info = json.loads(r.stderr.strip().splitlines()[-1])
        assert info["devices"] == (3 if dev else 1) and info["n"] == len(pos)
        outs[tag] = (read_dump(dump), open(surf, "rb").read())
    for a, b in zip(outs["one"][0], outs["three"][0]):
        assert np.array_equal(a["state"], b["state"])
    assert outs["one"][1] == outs["three"][1] and len(outs["one"][1]) > 8


@pytest.mark.gpu
def test_cli_generated_block_equals_xml_import(tmp_path):
    """`pbf_run --block nx ny nz` (the in-memory generator that stands in for a 10 GB XML file at 128M particles) builds the same
    particles, in the same order, as the XML import of the same lattice, on one device and on slabs."""
    import torch
    exe = _build()
    ng = torch.cuda.device_count()
    i, j, k = np.meshgrid(np.arange(30.0), np.arange(9.0), np.arange(8.0), indexing="ij")
    pos = np.stack([0.1 + 0.1 * i, 0.1 + 0.1 * j, 0.1 + 0.1 * k], axis=-1).reshape(-1, 3)
    vel = np.zeros_like(pos); vel[:, 1] = -1.0
    xml = str(tmp_path / "blk.xml")
    _write_xml(xml, pos, vel, 700.0)
    box = ["--box", "0", "0", "0", "4.5", "3.0", "1.1"]
    dumps = {}
    for tag, src in (("xml", ["-p", xml]), ("block", ["--block", "30", "9", "8", "--rho0", "700"]),
                     ("block2", ["--block", "30", "9", "8", "--devices", ",".join(str(d % ng) for d in range(2))])):
        dump = str(tmp_path / f"{tag}.bin")
        r = subprocess.run([exe] + src + ["--steps", "3", "--dump", dump, "--quiet"] + box, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        dumps[tag] = read_dump(dump)
    for a, b, c in zip(dumps["xml"], dumps["block"], dumps["block2"]):
        assert np.array_equal(a["state"], b["state"]) and np.array_equal(a["state"], c["state"])


@pytest.mark.gpu
def test_cli_lazy_mirror_equals_per_step_readback(tmp_path):
    """`--lazy-mirror` (Particles::mirror_each_step = false: one read-back after the last step) leaves the same final state and
    restart file as the default per-step read-back, on one device and on slabs."""
    import torch
    exe = _build()
    ng = torch.cuda.device_count()
    box = ["--box", "0", "0", "0", "4.5", "3.0", "1.1"]
    files = {}
    for tag, extra in (("eager", []), ("lazy", ["--lazy-mirror"]), ("lazy2", ["--lazy-mirror", "--devices", ",".join(str(d % ng) for d in range(2))])):
        ck = str(tmp_path / f"{tag}.ckpt")
        r = subprocess.run([exe, "--block", "30", "9", "8", "--steps", "4", "--save-state", ck, "--quiet"] + box + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        files[tag] = open(ck, "rb").read()
    assert files["eager"] == files["lazy"] and len(files["eager"]) > 30 * 9 * 8 * 48
    # the slab run stores the same particles (the checkpoint header records the device list only in memory, not in the file)
    assert files["eager"] == files["lazy2"]
