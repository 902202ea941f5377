set -x
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q ) > gpurun_out/r1_pytest_slab2.log 2>&1; tail -8 gpurun_out/r1_pytest_slab2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r1_bench_2gpu.json 2> gpurun_out/r1_bench_2gpu.err; tail -c 400 gpurun_out/r1_bench_2gpu.json; tail -3 gpurun_out/r1_bench_2gpu.err
