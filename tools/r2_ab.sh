#!/bin/bash
# generic A/B: bash tools/r2_ab.sh tag "ENV=.." "ENV2=.." ...   (first: GPU tests unless SKIP_PYTEST is set)
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -z "$SKIP_PYTEST" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log; fi
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 200 --warmup 10 ${BENCH_ARGS:---no-cpu-baseline} > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$i.json").read().strip().splitlines()[-1])
    ep=d.get("elbo_parity") or {}
    print("[$envs]:", round(d["value"],1), "it/s  e2e", round(d["e2e"]["value"],1), " parity", ep.get("ok"), ep.get("rel"), ep.get("mu_rel_fro"), ep.get("Sigma_rel_fro"))
    print("    phases", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["phases"].items() if v["ms_per_step"]>0})
except Exception as e:
    print("[$envs] failed", e); print(open("$OUT/bench_$i.err").read()[-1500:])
PY
done
