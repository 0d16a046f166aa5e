#!/bin/bash
export TAG=${1:-g7}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mosvgp or baseline_configs or softmax or multiclass or hetero or movgp or update_A" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -6 $OUT/pytest_sel.log
timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1.json 2> $OUT/bench_c5_n1.err; echo "bench c5 rc=$?"; tail -c 400 $OUT/bench_c5_n1.err
AGP_FAN=0 timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1_nofan.json 2> $OUT/bench_c5_n1_nofan.err; echo "bench c5 nofan rc=$?"
timeout 600 python bench.py --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"; tail -c 400 $OUT/bench_c4_n1.err
AGP_FAN=0 timeout 600 python bench.py --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n1_nofan.json 2> $OUT/bench_c4_n1_nofan.err; echo "bench c4 nofan rc=$?"
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
