set -x
timeout 600 python scripts/c3_1m.py > gpurun_out/r1_c3_1m.json 2> gpurun_out/r1_c3_1m.err; cat gpurun_out/r1_c3_1m.json; tail -3 gpurun_out/r1_c3_1m.err
