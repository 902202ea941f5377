#!/bin/bash
# ncu evidence for the solver kernels + launch list (run under gpurun)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 80 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py 400 200 200 1 > gpurun_out/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lambda|k_delta' -s 24 -c 2 -o gpurun_out/prof_ld python scripts/quick_bench.py 400 200 200 1 > gpurun_out/prof_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_build_neighbors' -s 1 -c 1 -o gpurun_out/prof_nb python scripts/quick_bench.py 400 200 200 1 >> gpurun_out/prof_run.log 2>&1
