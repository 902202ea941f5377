set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest_gpu.log; tail -5 gpurun_out/r1_pytest_gpu.log
QB_MESH=192 timeout 300 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_mesh.json 2>&1
QB_OBSTACLES=1 timeout 200 python scripts/quick_bench.py 400 200 200 3 > gpurun_out/r1_qb_obst.json 2>&1
grep -h "ms_per_step\"\|inside_mesh\|mesh:" gpurun_out/r1_qb_mesh.json gpurun_out/r1_qb_obst.json
