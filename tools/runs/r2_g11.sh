#!/bin/bash
export TAG=${1:-g11}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1.json 2> $OUT/bench_c5_n1.err; echo "bench c5 rc=$?"; tail -c 400 $OUT/bench_c5_n1.err
timeout 600 python bench.py --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"; tail -c 400 $OUT/bench_c4_n1.err
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
