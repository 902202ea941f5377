set -x
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r1_pytest_gpu.log 2>&1; tail -6 gpurun_out/r1_pytest_gpu.log
( time timeout 400 python bench.py ) > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; tail -c 200 gpurun_out/r1_bench_1gpu.json
