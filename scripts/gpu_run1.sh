set -x
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r1_pytest_gpu.log 2>&1; tail -6 gpurun_out/r1_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 400 python bench.py ) > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; tail -c 200 gpurun_out/r1_bench_1gpu.json
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r1_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r1_bench_reference.json
