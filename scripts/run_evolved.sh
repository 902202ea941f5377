mkdir -p gpurun_out; for z in 1 2 8; do PBF_ZSUB=$z python scripts/evolved_check.py > gpurun_out/evolved_z$z.json 2> gpurun_out/evolved_z$z.err; python - $z <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/evolved_z{sys.argv[1]}.json"))
print("zsub",sys.argv[1],"mismatch",d["digest_mismatch_total"],{k:(round(v["ke"],1),round(v["mean_nbrs"],1)) for k,v in d.items() if k.startswith("step")})
if sys.argv[1]=="1": print("ke64",[round(x,1) for x in d["ke_oracle64"]]); print("ke32",[round(x,1) for x in d["ke_oracle32"]])
PY
done
