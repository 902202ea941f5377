#!/bin/bash
# A/B of library builds under scripts/ubench/ (dev tool): same workload, per-kernel CUDA-event times
for v in "$@"; do
  PBF_LIB=$PWD/scripts/ubench/libpbf_$v.so python scripts/quick_bench.py 400 200 200 5 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{v}.json"))
    print(v, "ms/step %.2f"%d["ms_per_step"], "nbrs mean %.1f max %d"%(d["mean_nbrs"],d["max_nbrs"]), " ".join(f"{k}={x['ms_per_step']/max(x['launches_per_step'],1):.3f}" for k,x in d["kernels"].items() if x["ms_per_step"]>0.5))
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/ab_{v}.err").read()[-500:])
PY
done
