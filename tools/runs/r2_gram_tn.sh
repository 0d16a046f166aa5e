#!/bin/bash
# fused Gram kernel: harness first (correctness + timing), then parity tests + bench A/B if the harness is clean
OUT=gpurun_out/${1:-gram_tn}; mkdir -p $OUT
timeout 120 ./profiles/microbench/gemm_bench 8192 512 2>&1 | tee $OUT/gemm_bench_c2.txt
timeout 120 ./profiles/microbench/gemm_bench 16384 1024 2>&1 | tee $OUT/gemm_bench_c3.txt
timeout 120 ./profiles/microbench/gemm_bench 2048 256 2>&1 | tail -4 | tee $OUT/gemm_bench_small.txt
if grep -q "gram_tn launch: no error" $OUT/gemm_bench_c2.txt; then
  if [ -z "$SKIP_PYTEST" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log; fi
  for tn in 1 0; do
    AGP_GRAM_TN=$tn timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_tn$tn.json 2> $OUT/bench_tn$tn.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_tn$tn.json").read().strip().splitlines()[-1])
    print("gram_tn=$tn:", round(d["value"],1), "it/s  e2e", round(d["e2e"]["value"],1), " phases", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["phases"].items() if v["ms_per_step"]>0})
except Exception as e:
    print("gram_tn=$tn failed", e); print(open("$OUT/bench_tn$tn.err").read()[-1500:])
PY
  done
fi
