#!/bin/bash
export TAG=${1:-g10}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "--- gemm_bench ps"; timeout 60 ./profiles/microbench/gemm_bench 8192 512 > $OUT/gemm_bench_ps.txt 2>&1; echo "rc=$?"; cat $OUT/gemm_bench_ps.txt
echo "--- gemm_bench v1"; AGP_UMMA_PS=0 timeout 60 ./profiles/microbench/gemm_bench 8192 512 > $OUT/gemm_bench_v1.txt 2>&1; echo "rc=$?"; cat $OUT/gemm_bench_v1.txt
echo "--- gemm_bench ps C3"; timeout 60 ./profiles/microbench/gemm_bench 16384 1024 > $OUT/gemm_bench_ps_c3.txt 2>&1; echo "rc=$?"; cat $OUT/gemm_bench_ps_c3.txt
timeout 120 ./profiles/microbench/gemm_trace > $OUT/gemm_trace_ps.txt 2>&1; echo "trace rc=$?"
if grep -q "max abs err 1.1" $OUT/gemm_bench_ps.txt; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tf32x3 or full_size or baseline_configs or pipelined or hundred or predict" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -6 $OUT/pytest_sel.log
timeout 400 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_c2.err
timeout 600 python bench.py --config C3 --steps 50 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"; tail -c 400 $OUT/bench_c3.err
fi
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = (d.get("roofline") or {}).get("kernels") or {}
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"), "parity", (d.get("elbo_parity") or {}).get("ok"),
              {n: round(v["seconds_per_launch"] * 1e6, 1) for n, v in k.items() if "seconds_per_launch" in v})
    except Exception as e:
        print(f, "unreadable:", e)
PY
