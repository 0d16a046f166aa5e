#!/bin/bash
# full GPU suite + C3 / C4 / C5 single-GPU lines after the Gram-from-V change
OUT=gpurun_out/${1:-g14}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for cfg in C3 C4 C5; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 3 --no-cpu-baseline > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"],1), d["unit"], round(d["ms_per_step"]*1e3,1), "us/step  e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$cfg failed", e); print(open("$OUT/bench_$cfg.err").read()[-800:])
PY
done
