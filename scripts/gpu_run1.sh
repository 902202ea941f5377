set -x
timeout 600 python scripts/density_field_bench.py > gpurun_out/r1_density_surface.json 2> gpurun_out/r1_density_surface.err; tail -40 gpurun_out/r1_density_surface.json; tail -5 gpurun_out/r1_density_surface.err
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1_pytest_gpu.log
