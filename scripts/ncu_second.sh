#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_build_neighbors|k_vorticity_xsph|k_confine_commit|k_cell_sort' -s 4 -c 4 -o gpurun_out/prof_misc python scripts/quick_bench.py 400 200 200 1 > gpurun_out/prof_run2.log 2>&1
