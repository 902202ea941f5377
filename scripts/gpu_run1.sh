set -x
( time timeout 900 python -m pytest tests -m gpu -x -q -k "surface_extraction" ) > gpurun_out/r1_pytest_surface.log 2>&1; tail -30 gpurun_out/r1_pytest_surface.log
