set -x
( time timeout 400 python bench.py ) > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; tail -c 300 gpurun_out/r1_bench_1gpu.json; tail -5 gpurun_out/r1_bench_1gpu.err
