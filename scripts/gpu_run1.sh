set -x
python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1; tail -3 gpurun_out/r1_smoke.log
( timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py ) > gpurun_out/r1_sanitizer.txt 2>&1; tail -6 gpurun_out/r1_sanitizer.txt
