#!/bin/bash
# round-2 GPU check: usage  gpurun --timeout 1500 -- 'bash tools/r2_check.sh <tag> [stage ...]'   stages: tests bench ab c5 c4 c3
export TAG=${1:-r2}; shift
STAGES=${*:-tests bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has quick; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tf32x3_parity or full_size or baseline_configs or pipelined_pool or mosvgp_parity" > $OUT/pytest_quick.log 2>&1; echo "quick pytest rc=$?"; tail -5 $OUT/pytest_quick.log
fi
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
fi
if has bench; then
  timeout 400 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; tail -c 600 $OUT/bench_c2.err
fi
if has ab; then
  AGP_TAIL_VARIANT=2 timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_c2_tail2.json 2> $OUT/bench_c2_tail2.err; echo "bench tail2 rc=$?"
fi
if has c5; then
  timeout 600 python bench.py --config C5 --steps 20 --warmup 3 > $OUT/bench_c5_n1.json 2> $OUT/bench_c5_n1.err; echo "bench c5 rc=$?"; tail -c 600 $OUT/bench_c5_n1.err
fi
if has c4; then
  timeout 600 python bench.py --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"; tail -c 600 $OUT/bench_c4_n1.err
fi
if has c3; then
  timeout 600 python bench.py --config C3 --steps 50 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"; tail -c 600 $OUT/bench_c3.err
fi
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = (d.get("roofline") or {}).get("kernels") or {}
        print(os.path.basename(f), round(d["value"]), "latent-it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"),
              {n: round(v["seconds_per_launch"] * 1e6, 1) for n, v in k.items() if "seconds_per_launch" in v}, "parity", d.get("elbo_parity"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
