#!/bin/bash
# quick GPU check: parity tests + bench line(s).  usage: bash tools/gpu_quick.sh tag ["ENV=.. ENV2=.."]...
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$i.json"))
    print("[$envs]:", round(d["value"],1), "it/s  e2e", round(d["e2e"]["value"],1), " chol", round(d["phases"]["chol_blocked"]["ms_per_step"]*1e3,1), "us  elbo", d["elbo_last"])
except Exception as e:
    print("[$envs] failed", e); print(open("$OUT/bench_$i.err").read()[-2000:])
PY
done
