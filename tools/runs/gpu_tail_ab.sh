#!/bin/bash
# A/B of the m x m tail variants: microbenchmarks + the bench line per AGP_TAIL_VARIANT + parity tests on the new variant
OUT=gpurun_out/${1:-tail_ab}
mkdir -p $OUT
./profiles/microbench/dp_latency > $OUT/dp_latency.txt 2>&1; cat $OUT/dp_latency.txt
./profiles/microbench/potf2_bench > $OUT/potf2_bench.txt 2>&1; cat $OUT/potf2_bench.txt
for v in 0 1 2; do
  AGP_TAIL_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_v$v.json 2> $OUT/bench_v$v.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_v$v.json"))
print("variant $v:", round(d["value"],1), "it/s  chol", round(d["phases"]["chol_blocked"]["ms_per_step"]*1e3,1), "us  elbo", d["elbo_last"])
PY
done
AGP_TAIL_VARIANT=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
