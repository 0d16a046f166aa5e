#!/bin/bash
export TAG=${1:-g6}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined_pool or tf32x3 or full_size or hundred or step_batch_async" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -12 $OUT/pytest_sel.log
AGP_EARLY_STATS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined_pool" > $OUT/pytest_early.log 2>&1; echo "early pytest rc=$?"; tail -4 $OUT/pytest_early.log
timeout 400 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_c2.err
AGP_SPLIT_GRAM=0 timeout 400 python bench.py --no-cpu-baseline > $OUT/bench_c2_nosplit.json 2> $OUT/bench_c2_nosplit.err; echo "bench nosplit rc=$?"
AGP_EARLY_STATS=1 timeout 400 python bench.py --no-cpu-baseline > $OUT/bench_c2_early.json 2> $OUT/bench_c2_early.err; echo "bench early rc=$?"
timeout 600 python bench.py --config C3 --steps 50 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"; tail -c 400 $OUT/bench_c3.err
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"), "parity", (d.get("elbo_parity") or {}).get("ok"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
