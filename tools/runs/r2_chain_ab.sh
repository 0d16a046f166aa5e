#!/bin/bash
# A/B of the pivot-chain variants inside the engine (same box, interleaved).  usage: bash tools/r2_chain_ab.sh tag
TAG=${1:-chain_ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do
  for c in 0 1; do
    AGP_CHAIN=$c timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_chain${c}_$rep.json 2> $OUT/bench_chain${c}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_chain${c}_$rep.json").read().strip().splitlines()[-1])
    print("chain=$c rep=$rep:", round(d["value"],1), "it/s  e2e", round(d["e2e"]["value"],1), " chol", round(d["phases"]["chol_blocked"]["ms_per_step"]*1e3,1), "us  parity", d["elbo_parity"]["ok"], d["elbo_parity"]["Sigma_rel_fro"])
except Exception as e:
    print("chain=$c failed", e); print(open("$OUT/bench_chain${c}_$rep.err").read()[-1500:])
PY
  done
done
