set -x
PBF_TEX=0 timeout 200 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_tex0.json 2>&1
PBF_TEX=1 timeout 200 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_tex1.json 2>&1
grep -h -A1 "\"lambda\"\|\"ms_per_step\":\|avg_rho" gpurun_out/r1_qb_tex0.json gpurun_out/r1_qb_tex1.json | grep -v "^--"
