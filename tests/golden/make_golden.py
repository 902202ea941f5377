#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

Needs /root/reference and oracle/_ref/ref_harness (make -C oracle ref).  The GPU box has neither;
it only reads the committed .npz files.  Re-run after changing the harness:  python tests/golden/make_golden.py

Fixtures
  scene_<name>.npz        inputs of the reference's shipped scenes (particles/p.xml, particles/spheres_p.xml)
                          parsed like Application::load_particles (application.cpp:302-344)
  ref_<name>.npz          reference output for 20 steps of timeStep(): sha256[:16] of <f8[N,7] per step,
                          coordinate sums, directed pair counts, the "avg rho: a => b" lines the reference
                          prints, full state at steps 0/1/19 and the ordered neighbour CSR of step 0
  ref_sphere_<name>.npz   same with the CBspheres obstacle spheres in the reference's BVH (harness --sphere)
  ref_mesh_drop.npz       same with obstacle triangles (a cuboid and a wedge) in the reference's BVH (harness --tris)
  ref_surface_<name>.npz  Particles::getSurfacePrims(0.95 rho0, 0.15) (= updateSurface, particles.cpp:393-402) of the
                          unmodified reference after a few steps: the state and the triangle soup (18 doubles each)
  mc_cases.npz            the reference's polygonise() (marching.cpp:17) on the unit cell for all 256 sign patterns
  ref_jitter_<name>.npz   same for jittered pgen-style inputs (lattice + U(-0.001,0.001),
                          default_rng(1234)) — the inputs used for fp32-vs-fp64 tolerance checks
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import (GOLDEN, REFERENCE, box_mesh, load_xml_scene, pgen_two_blocks, lattice_block, ramp_mesh, read_dump,
                     ref_harness_path, state_sha, write_bin_scene, write_tris)

STEPS = 20
KEEP = (0, 1, 19)


def run_reference(pos, vel, rho0, steps, spheres=(), tris=None):
    with tempfile.TemporaryDirectory() as td:
        scene = os.path.join(td, "scene.bin"); dump = os.path.join(td, "dump.bin")
        write_bin_scene(scene, pos, vel, rho0)
        cmd = [ref_harness_path(), "--bin", scene, "--steps", str(steps), "--out", dump, "--quiet"]
        for sp in spheres:                      # StaticScene::Sphere primitives added to the reference's BVH
            cmd += ["--sphere"] + [repr(float(x)) for x in sp]
        if tris is not None:                    # MarchingTriangle primitives added after the spheres
            tf = os.path.join(td, "tris.bin"); write_tris(tf, tris); cmd += ["--tris", tf]
        subprocess.run(cmd, check=True)
        log = open(dump + ".log").read()
        return read_dump(dump), log


def density_queries():
    """Marching-cubes style lattice over the Cornell box at step H/2 (FSTEPSIZE_RATIO 0.5,
    particles.cpp:18) plus random points, some outside the box."""
    g = np.stack(np.meshgrid(np.arange(-1.05, 1.06, 0.15), np.arange(-0.05, 1.56, 0.15), np.arange(-1.05, 1.06, 0.15), indexing="ij"), -1).reshape(-1, 3)
    r = np.random.default_rng(77).uniform([-1.3, -0.3, -1.3], [1.3, 1.8, 1.3], size=(500, 3))
    return np.concatenate([g, r])


def reference_density_field(pos, vel, rho0, q):
    """Particles::estimateDensityAt (particles.cpp:446-453) of the unmodified reference on the loaded scene."""
    with tempfile.TemporaryDirectory() as td:
        scene = os.path.join(td, "scene.bin"); qf = os.path.join(td, "q.bin"); df = os.path.join(td, "d.bin")
        write_bin_scene(scene, pos, vel, rho0)
        with open(qf, "wb") as f:
            f.write(np.int64(q.shape[0]).tobytes()); f.write(np.ascontiguousarray(q, dtype=np.float64).tobytes())
        subprocess.run([ref_harness_path(), "--bin", scene, "--steps", "0", "--density-queries", qf, "--density-out", df, "--quiet"], check=True)
        return np.fromfile(df, dtype=np.float64)


def reference_surface(pos, vel, rho0, steps):
    """State after `steps` reference steps and the reference's own marching-cubes surface of it."""
    with tempfile.TemporaryDirectory() as td:
        scene = os.path.join(td, "scene.bin"); dump = os.path.join(td, "dump.bin"); sf = os.path.join(td, "surf.bin")
        write_bin_scene(scene, pos, vel, rho0)
        subprocess.run([ref_harness_path(), "--bin", scene, "--steps", str(steps), "--out", dump, "--surface", sf, "--quiet"], check=True)
        d = read_dump(dump)
        raw = np.fromfile(sf, dtype=np.uint8)
        nt = int(raw[:8].view(np.int64)[0])
        tris = raw[8:].view(np.float64).reshape(nt, 18).copy()
        return d[-1]["state"].copy(), tris


def reference_mc_cases():
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "cases.bin")
        subprocess.run([ref_harness_path(), "--mc-cases", f], check=True)
        raw = np.fromfile(f, dtype=np.uint8)
    counts, verts, off = [], [], 0
    for c in range(256):
        nt = int(raw[off:off + 8].view(np.int64)[0]); off += 8
        v = raw[off:off + 72 * nt].view(np.float64).reshape(nt, 9); off += 72 * nt
        counts.append(nt); verts.append(v)
    assert off == len(raw)
    return np.array(counts), np.concatenate(verts)


def pack(dump, log, keep):
    lines = re.findall(r"avg rho: (\S+) => (\S+)", log)
    out = dict(
        sha=np.array([state_sha(d["state"][:, 0:3], d["state"][:, 3:6], d["state"][:, 6]) for d in dump]),
        sums=np.array([d["state"][:, 0:3].sum(axis=0) for d in dump]),
        pairs=np.array([len(d["col"]) for d in dump], dtype=np.int64),
        avg_rho_text=np.array(lines),
        keep=np.array(keep, dtype=np.int64),
    )
    for s in keep:
        out[f"state_{s}"] = dump[s]["state"][:, :7].copy()
    out["nbr_counts_0"] = dump[0]["counts"].copy()
    out["nbr_col_0"] = dump[0]["col"].copy()
    return out


def jitter_scenes():
    """name -> (pos, vel, rho0); all inside the reference's hard-coded Cornell box."""
    pos, vel, rho0 = pgen_two_blocks()
    rng = np.random.default_rng(1234)
    two = pos + rng.uniform(-0.001, 0.001, size=pos.shape)
    # p.xml-like sparse block: 12x5x12, spacing 0.15
    pb, vb = lattice_block(12, 5, 12, origin=(-0.9, 0.5, -0.9), spacing=0.15, v0=(0, -0.01, 0), jitter=0.0015, seed=1234)
    # block resting on the floor against two walls: exercises floor/wall contacts and the slide
    pc, vc = lattice_block(8, 10, 8, origin=(-0.97, 0.03, -0.97), spacing=0.1, v0=(-0.5, -1.0, -0.25), jitter=0.001, seed=99)
    # block thrown at the open front / light planes (virtual planes z=1, y=1.49)
    pd, vd = lattice_block(6, 6, 6, origin=(0.2, 0.95, 0.45), spacing=0.1, v0=(0.3, 2.5, 3.0), jitter=0.001, seed=7)
    return {
        "two_blocks": (two, vel, rho0),
        "sparse": (pb, vb, 150.0),
        "corner": (pc, vc, 700.0),
        "front": (pd, vd, 700.0),
    }


# the two r = 0.3 obstacle spheres of the CBspheres scenes (dae/sky/CBspheres_lambertian.dae:291-305,575-594;
# the file is z-up: (x, y, z) -> (x, z, -y)), both resting on the floor
CB_SPHERES = np.array([[-0.4, 0.3, -0.3, 0.3], [0.4, 0.3, 0.3, 0.3]])


def sphere_scenes():
    """name -> (pos, vel, rho0, spheres, steps, keep): obstacle-sphere collision."""
    two, vel, rho0 = jitter_scenes()["two_blocks"]
    # "sphere drop" (BASELINE config 2): spheres_p.xml-style blocks collapsing around the CBspheres spheres
    # a block thrown straight down onto sphere 1: direct hits and slides from the second step on
    ph, vh = lattice_block(8, 8, 8, origin=(-0.75, 0.65, -0.65), spacing=0.1, v0=(0.0, -3.0, 0.0), jitter=0.001, seed=21)
    return {
        "sphere_drop": (two, vel, rho0, CB_SPHERES, 40, (0, 1, 19, 20, 39)),
        "sphere_hit": (ph, vh, 700.0, CB_SPHERES, 12, (0, 1, 2, 3, 4, 5, 11)),
    }


def mesh_scene():
    """Obstacle triangles: the two collapsing blocks around a cuboid standing on the floor
    (10 triangles, geometric normals) and a wedge with averaged, non-unit vertex normals along its top edge and one
    clockwise triangle (6 triangles) in the two empty quadrants."""
    two, vel, rho0 = jitter_scenes()["two_blocks"]
    tris = np.concatenate([box_mesh((-0.7, 0.0, -0.7), (-0.2, 0.4, -0.2)), ramp_mesh(0.15, 0.9, 0.5, 0.2, 0.8)])
    return two, vel, rho0, tris, 40, (0, 1, 9, 10, 11, 12, 13, 19, 20, 39)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    pos, vel, rho0, tris, steps, keep = mesh_scene()
    dump, log = run_reference(pos, vel, rho0, steps, tris=tris)
    np.savez_compressed(os.path.join(GOLDEN, "ref_mesh_drop.npz"), pos=pos, vel=vel, rho0=rho0, tris=tris, **pack(dump, log, keep))
    print("mesh_drop", pos.shape[0], "pairs", [len(d["col"]) for d in dump][:6])
    if "--only-mesh" in sys.argv:
        return
    for name, (pos, vel, rho0, sph, steps, keep) in sphere_scenes().items():
        dump, log = run_reference(pos, vel, rho0, steps, sph)
        np.savez_compressed(os.path.join(GOLDEN, f"ref_{name}.npz"), pos=pos, vel=vel, rho0=rho0, spheres=sph, **pack(dump, log, keep))
        inside = [int((np.linalg.norm(dump[-1]["state"][:, 0:3] - c[:3], axis=1) < c[3]).sum()) for c in sph]
        print("spheres", name, pos.shape[0], "pairs", [len(d["col"]) for d in dump][:6], "particles inside a sphere at the end:", inside)
    if "--only-spheres" in sys.argv:
        return
    for name in ("p", "spheres_p"):
        pos, vel, rho0 = load_xml_scene(os.path.join(REFERENCE, "particles", name + ".xml"))
        np.savez_compressed(os.path.join(GOLDEN, f"scene_{name}.npz"), pos=pos, vel=vel, rho0=rho0)
        dump, log = run_reference(pos, vel, rho0, STEPS)
        np.savez_compressed(os.path.join(GOLDEN, f"ref_{name}.npz"), **pack(dump, log, KEEP))
        print(name, pos.shape[0], "sha[0,1,19] =", [state_sha(dump[s]["state"][:, 0:3], dump[s]["state"][:, 3:6], dump[s]["state"][:, 6]) for s in KEEP])
    for name, (pos, vel, rho0) in jitter_scenes().items():
        dump, log = run_reference(pos, vel, rho0, 6)
        np.savez_compressed(os.path.join(GOLDEN, f"ref_jitter_{name}.npz"), pos=pos, vel=vel, rho0=rho0, **pack(dump, log, (0, 1, 5)))
        print("jitter", name, pos.shape[0], "pairs", [len(d["col"]) for d in dump])
    # density field (marching-cubes input) of two loaded scenes
    q = density_queries()
    for name, (pos, vel, rho0) in list(jitter_scenes().items())[:2]:
        d = reference_density_field(pos, vel, rho0, q)
        np.savez_compressed(os.path.join(GOLDEN, f"ref_density_{name}.npz"), pos=pos, vel=vel, rho0=rho0, q=q, density=d)
        print("density field", name, q.shape[0], "points, max", d.max())


def surface_fixtures():
    for name, steps in (("p", 12), ("spheres_p", 30)):
        sc = np.load(os.path.join(GOLDEN, f"scene_{name}.npz"))
        st, tris = reference_surface(sc["pos"], sc["vel"], float(sc["rho0"]), steps)
        np.savez_compressed(os.path.join(GOLDEN, f"ref_surface_{name}.npz"), state=st, tris=tris, rho0=float(sc["rho0"]), steps=steps)
        print(f"ref_surface_{name}: {len(tris)} triangles after {steps} steps")
    counts, verts = reference_mc_cases()
    np.savez_compressed(os.path.join(GOLDEN, "mc_cases.npz"), counts=counts, verts=verts)
    print(f"mc_cases: {counts.sum()} triangles over 256 patterns")


if __name__ == "__main__":
    if "--surface-only" in sys.argv:
        surface_fixtures(); sys.exit(0)
    main()
