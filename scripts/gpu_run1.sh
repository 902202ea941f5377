set -x
ncu --set full --clock-control none --import-source on -k regex:k_build_neighbors -s 3 -c 1 -o gpurun_out/prof_nb_uniform -f python scripts/quick_bench.py 400 200 200 2 > gpurun_out/ncu_nb.log 2>&1; tail -3 gpurun_out/ncu_nb.log
PBF_NB_PER_LANE=1 ncu --set full --clock-control none --import-source on -k regex:k_build_neighbors -s 3 -c 1 -o gpurun_out/prof_nb_lane -f python scripts/quick_bench.py 400 200 200 2 > gpurun_out/ncu_nb2.log 2>&1; tail -3 gpurun_out/ncu_nb2.log
ls -la gpurun_out/*.ncu-rep
