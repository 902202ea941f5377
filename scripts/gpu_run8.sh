set -x
N=$1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r1_bench_${N}gpu.json 2> gpurun_out/r1_bench_${N}gpu.err; tail -c 300 gpurun_out/r1_bench_${N}gpu.json; tail -3 gpurun_out/r1_bench_${N}gpu.err
