#!/bin/bash
# ncu evidence taken from bench.py itself (run under gpurun, 1 GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
# launch list: every launch of the timed region with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 109 -c 72 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu_bench_run.log 2>&1
# full sets for the solver kernels and the neighbour build / finalize kernels
ncu --set full --clock-control none --import-source on -k regex:'k_lambda|k_delta' -s 48 -c 2 -o gpurun_out/prof_bench_ld $BENCH >> gpurun_out/ncu_bench_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_build_neighbors|k_vorticity_xsph|k_confine_commit|k_cell_sort' -s 4 -c 4 -o gpurun_out/prof_bench_misc $BENCH >> gpurun_out/ncu_bench_run.log 2>&1
