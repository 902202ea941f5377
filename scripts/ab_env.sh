#!/bin/bash
# A/B of one library under different environment settings (dev tool): ab_env.sh <lib tag> VAR=a VAR=b ...
v=$1; shift
for kv in "$@"; do
  env $kv PBF_LIB=$PWD/scripts/ubench/libpbf_$v.so python scripts/quick_bench.py 400 200 200 5 > gpurun_out/ab_${v}_$kv.json 2> gpurun_out/ab_${v}_$kv.err
  python - "$v" "$kv" <<'PY'
import json,sys
v,kv=sys.argv[1:3]
try:
    d=json.load(open(f"gpurun_out/ab_{v}_{kv}.json"))
    print(v, kv, "ms/step %.2f"%d["ms_per_step"], "nbrs mean %.1f max %d"%(d["mean_nbrs"],d["max_nbrs"]), " ".join(f"{k}={x['ms_per_step']/max(x['launches_per_step'],1):.3f}" for k,x in d["kernels"].items() if x["ms_per_step"]>0.3))
except Exception as e:
    print(v, kv, "failed", e, open(f"gpurun_out/ab_{v}_{kv}.err").read()[-500:])
PY
done
