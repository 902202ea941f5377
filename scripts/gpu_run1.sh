set -x
( time timeout 900 python -m pytest tests -m gpu -x -q -k "edge_cases or obstacle_mesh" ) > gpurun_out/r1_pytest_new.log 2>&1; tail -30 gpurun_out/r1_pytest_new.log
