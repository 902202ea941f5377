#!/bin/bash
# ncu --set full of k_lambda and k_delta on the SETTLED state of the bench workload (C4 after 150 + 3 steps), run under gpurun, 1 GPU.
# 153 steps x 12 launches of each kernel are skipped, the next launch of each is captured.
mkdir -p gpurun_out
cat > /tmp/evolved_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from fluid_b200 import api
nx, ny, nz = 400, 200, 200
prm = api.default_params(rest_density=700.0, iterations=12, box_min=(0, 0, 0), box_max=(120.0, 30.0, 20.1), y_light=30.0, z_front=20.1)
g = api.Solver(prm)
pos, vel = bench.block_f32(nx, ny, nz)
g.upload(pos, vel); g.step(155)
d, c = g.neighbor_digest()
print("mean neighbours", float(c.mean()), "max", int(c.max()))
PY
ncu --set full --clock-control none --import-source on -k regex:'k_lambda|k_delta' -s 3672 -c 2 -o gpurun_out/prof_evolved_ld python /tmp/evolved_run.py > gpurun_out/ncu_evolved_run.log 2>&1
tail -3 gpurun_out/ncu_evolved_run.log
