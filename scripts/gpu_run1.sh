set -x
PBF_NB_PER_LANE=0 timeout 200 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_nb_uni2.json 2>&1
grep -A2 '"build_neighbors"' gpurun_out/r1_qb_nb_uni2.json; grep '"ms_per_step":\|mean_nbrs' gpurun_out/r1_qb_nb_uni2.json | grep -v "   "
( time timeout 900 python -m pytest tests -m gpu -x -q -k "neighbor or crowded or capacity or large_block" ) > gpurun_out/r1_pytest_nb.log 2>&1; tail -3 gpurun_out/r1_pytest_nb.log
