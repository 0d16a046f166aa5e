#!/bin/bash
# usage: gpurun --timeout 1200 -- 'bash tools/r2_e2e.sh <tag>'   -- asynchronous host-batch path + grouped multi-latent launches: parity tests + benches
export TAG=${1:-e2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_batch_async or mosvgp or baseline_configs or softmax or multiclass or hetero" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -15 $OUT/pytest_sel.log
timeout 400 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_c2.err
AGP_ASYNC_SERIAL=1 timeout 400 python bench.py --no-cpu-baseline > $OUT/bench_c2_serial.json 2> $OUT/bench_c2_serial.err; echo "bench serial rc=$?"
timeout 400 python bench.py --no-cpu-baseline --e2e-dtype f32 > $OUT/bench_c2_f32rows.json 2> $OUT/bench_c2_f32rows.err; echo "bench f32 rows rc=$?"
timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1.json 2> $OUT/bench_c5_n1.err; echo "bench c5 rc=$?"; tail -c 400 $OUT/bench_c5_n1.err
AGP_NO_GROUPED=1 timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1_nogroup.json 2> $OUT/bench_c5_n1_nogroup.err; echo "bench c5 nogroup rc=$?"
timeout 600 python bench.py --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"; tail -c 400 $OUT/bench_c4_n1.err
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"), "parity", (d.get("elbo_parity") or {}).get("ok"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
