#!/bin/bash
# A/B of environment settings on the same library (dev tool): scripts/ab_env2.sh name "ENV=.. ENV=.." name2 "..."
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  env $envs python scripts/quick_bench.py 400 200 200 5 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{v}.json"))
    print(v, "ms/step %.2f"%d["ms_per_step"], "avg_rho", d["avg_rho"], " ".join(f"{k}={x['ms_per_step']/max(x['launches_per_step'],1):.3f}" for k,x in d["kernels"].items() if x["ms_per_step"]>0.5))
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/ab_{v}.err").read()[-500:])
PY
done
