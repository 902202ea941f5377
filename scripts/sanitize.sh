#!/bin/bash
# compute-sanitizer passes (run under gpurun, 1 GPU); logs -> gpurun_out/sanitize_<tool>[_multi].log
#   scripts/sanitize.sh [memcheck racecheck synccheck]          scripts/sanitize_small.py  (single-GPU paths)
#   SCRIPT=scripts/sanitize_multi.py TAG=_multi scripts/sanitize.sh memcheck synccheck     (three peer-mode slabs on one GPU)
mkdir -p gpurun_out
SCRIPT=${SCRIPT:-scripts/sanitize_small.py}
for tool in ${@:-memcheck racecheck synccheck}; do
  echo "+ compute-sanitizer --tool $tool python $SCRIPT" > gpurun_out/sanitize_$tool$TAG.log
  timeout ${LIMIT:-1500} compute-sanitizer --tool $tool python $SCRIPT >> gpurun_out/sanitize_$tool$TAG.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool$TAG.log
  tail -4 gpurun_out/sanitize_$tool$TAG.log
done
