#!/bin/bash
# compute-sanitizer passes over scripts/sanitize_small.py (run under gpurun, 1 GPU); logs -> gpurun_out/sanitize_<tool>.log
mkdir -p gpurun_out
for tool in ${@:-memcheck racecheck synccheck}; do
  echo "+ compute-sanitizer --tool $tool python scripts/sanitize_small.py" > gpurun_out/sanitize_$tool.log
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_small.py >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  tail -4 gpurun_out/sanitize_$tool.log
done
