"""Marching-cubes density field (Particles::estimateDensityAt at every lattice vertex, step H/2,
particles.cpp:332-348,446-453) on the GPU vs the CPU restatement.  Dev/report tool."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from fluid_b200 import api
from bench import block_f32

out = []
for dims in ((64, 32, 32), (400, 200, 200)):
    pos, vel = block_f32(*dims)
    n = len(pos)
    box_max = (max(30.0, 0.3 * dims[0]), 30.0, 0.1 * dims[2] + 0.1)
    g = api.Solver(api.default_params(rest_density=700.0, box_min=(0, 0, 0), box_max=box_max, y_light=box_max[1], z_front=box_max[2]))
    g.upload(pos, vel); g.step(2)
    P, _, _ = g.download()
    lo, hi = P.min(0) - 0.3, P.max(0) + 0.3
    ax = [np.arange(lo[a], hi[a], 0.15) for a in range(3)]
    q = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    g.density_at(q[:1000])                       # re-bin + warm-up
    t0 = time.perf_counter(); d = g.density_at(q); dt = time.perf_counter() - t0
    row = {"particles": n, "lattice_points": len(q), "gpu_seconds_incl_transfers": dt, "gpu_points_per_s": len(q) / dt}
    if n <= 100000:
        o = H.Oracle(H.default_params(rest_density=700.0), 64, H.COLLIDE_BOX, H.SEARCH_GRID); o.upload(P, np.zeros_like(P))
        t0 = time.perf_counter(); do = o.density_at(q[:20000]); dtc = time.perf_counter() - t0
        row.update({"cpu_reference_algorithm_points_per_s": 20000 / dtc, "cpu_threads": H.oracle_lib().oracle_max_threads(),
                    "max_rel_err_vs_cpu": float(np.abs(d[:20000] - do).max() / do.max())})
    # the whole surfacer on the device (pbf_extract_surface): lattice over the fluid's bounding box, step H/2, iso 0.95 rho0
    slo, shi = tuple(float(v) for v in (P.min(0) - 0.3)), tuple(float(v) for v in (P.max(0) + 0.3))
    g.extract_surface(700.0, lo=slo, hi=(slo[0] + 1.0, slo[1] + 1.0, slo[2] + 1.0))      # warm-up (table upload)
    t0 = time.perf_counter(); tris = g.extract_surface(700.0, lo=slo, hi=shi); dts = time.perf_counter() - t0
    cells = int(np.prod([int((shi[a] - slo[a]) / 0.15) + 1 for a in range(3)]))
    row.update({"surface_lattice_cells": cells, "surface_triangles": len(tris), "surface_gpu_seconds_incl_readback": dts,
                "surface_density_evaluations": 8 * cells + 18 * len(tris)})
    if n <= 100000:
        t0 = time.perf_counter(); ts = o.surface(700.0, lo=slo, hi=shi); dtc = time.perf_counter() - t0
        row.update({"surface_cpu_reference_algorithm_seconds": dtc, "surface_same_triangle_count": bool(len(ts) == len(tris)),
                    "surface_max_abs_diff_vs_cpu": float(np.abs(ts - tris).max()) if len(ts) == len(tris) else None})
    out.append(row)
print(json.dumps(out, indent=1))
