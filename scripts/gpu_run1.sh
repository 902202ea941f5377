set -x
( time timeout 600 python bench.py --block 3200 200 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/r1_bench_1gpu_128m.json 2> gpurun_out/r1_bench_1gpu_128m.err; tail -c 300 gpurun_out/r1_bench_1gpu_128m.json; tail -5 gpurun_out/r1_bench_1gpu_128m.err
