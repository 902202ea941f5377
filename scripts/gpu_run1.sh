set -x
PBF_SMEM=1 timeout 200 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_smem2.json 2> gpurun_out/r1_qb_smem2.err
grep -A2 '"lambda"' gpurun_out/r1_qb_smem2.json; grep PBF_SMEM gpurun_out/r1_qb_smem2.err; grep avg_rho -A2 gpurun_out/r1_qb_smem2.json
