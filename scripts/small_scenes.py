"""BASELINE.json configs C1/C2 (the reference's shipped scenes): GPU step time vs the unmodified
reference on the host (oracle/_ref/ref_harness) and the oracle port.  Dev/report tool."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api

out = []
for name, iters, steps in (("p", 12, 63), ("spheres_p", 4, 63), ("spheres_p", 12, 63)):
    sc = np.load(os.path.join(H.GOLDEN, f"scene_{name}.npz"))
    pos, vel, rho0 = sc["pos"], sc["vel"], float(sc["rho0"])
    n = len(pos)
    row = {"scene": name, "n": n, "iterations": iters, "steps": steps}
    for graph in ("0", "1"):
        os.environ["PBF_GRAPH"] = graph          # 1 (the default for small scenes) = replay the step's launches as one CUDA graph
        g = api.Solver(api.default_params(rest_density=rho0, iterations=iters))
        g.upload(pos, vel); g.step(5)
        g.upload(pos, vel)
        t0 = time.perf_counter(); g.step(steps); wall = time.perf_counter() - t0
        ms = g.stats()[2]
        # the host adapter's pattern: one step + read-back per frame (Particles::timeStep, then redraw reads ps)
        t0 = time.perf_counter()
        for _ in range(steps):
            g.step(1); g.download()
        row[f"gpu_wall_ms_per_step_with_download_graph{graph}"] = 1e3 * (time.perf_counter() - t0) / steps
        row[f"gpu_ms_per_step_graph{graph}"] = ms / steps
        row[f"gpu_wall_ms_per_step_graph{graph}"] = 1e3 * wall / steps
        row[f"gpu_updates_per_s_graph{graph}"] = n * iters * steps / (ms * 1e-3)
    # CPU: oracle port (fp64, grid, all cores) and, for 12 iterations, the unmodified reference
    o = H.Oracle(H.default_params(rest_density=rho0, iterations=iters, xsph_mode=H.XSPH_JACOBI), 64, H.COLLIDE_BOX, H.SEARCH_GRID)
    o.upload(pos, vel); t0 = time.perf_counter(); o.step(steps); dt = time.perf_counter() - t0
    row["oracle_port_ms_per_step"] = 1e3 * dt / steps; row["oracle_threads"] = H.oracle_lib().oracle_max_threads()
    if iters == 12 and H.have_reference_binary():
        for opt in ("O3", "O0"):
            scene = f"/tmp/{name}.bin"; dump = f"/tmp/{name}.dump"
            H.write_bin_scene(scene, pos, vel, rho0)
            subprocess.run([H.ref_harness_path(opt), "--bin", scene, "--steps", str(steps), "--out", dump, "--quiet"], check=True)
            secs = [d["seconds"] for d in H.read_dump(dump)]
            row[f"reference_{opt}_ms_per_step"] = 1e3 * sum(secs) / len(secs)
            row[f"reference_{opt}_updates_per_s"] = n * iters * len(secs) / sum(secs)
    out.append(row)
print(json.dumps(out, indent=1))
