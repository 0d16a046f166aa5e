#!/bin/bash
export TAG=${1:-e2e2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_batch_async" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -5 $OUT/pytest_sel.log
timeout 400 python bench.py --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_c2.err
timeout 400 python bench.py --no-cpu-baseline --e2e-dtype f32 > $OUT/bench_c2_f32rows.json 2> $OUT/bench_c2_f32rows.err; echo "bench f32 rows rc=$?"
python - <<'PY'
import torch, time
x = torch.empty(8192*32, dtype=torch.float64).pin_memory(); d = torch.empty_like(x, device="cuda")
for n in (1, 4):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20 * n): d.copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D 2.1 MB pinned: %.1f us per copy, %.1f GB/s" % (e0.elapsed_time(e1) * 1e3 / (20 * n), x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3 / (20 * n)) / 1e9))
PY
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "") or "*", "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", d.get("e2e", {}).get("value"), "parity", (d.get("elbo_parity") or {}).get("ok"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
