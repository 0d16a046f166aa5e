#!/bin/bash
# First GPU contact of the experimental code written without a GPU.  Everything runs under
# short timeouts; the new kernel's mbarrier waits trap after ~2 s instead of hanging.
# usage: gpurun --timeout 900 -- 'bash tools/umma_v2_check.sh [tag] [stage ...]'      stages: gemm ns async   (default: all)
#   AGP_UMMA_V2=1  v2 GEMM kernel, L^-1 / X pre-split        AGP_UMMA_V2=3  v2 kernel, right operand split inside the kernel
export TAG=${1:-umma_v2}
shift
STAGES=${*:-gemm ns async}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has gemm; then
  for V in 1 3; do
    # the tf32x3 parity tests exercise all three GEMM launches (V, V X^T statistics, Gram)
    AGP_UMMA_V2=$V timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tf32x3 or full_size or baseline_configs or pipelined_pool or ragged_sizes or knm_tensor_core" \
        > $OUT/pytest_v2_$V.log 2>&1; echo "v2=$V pytest rc=$?" | tee -a $OUT/pytest_v2_$V.log
    tail -4 $OUT/pytest_v2_$V.log
  done
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_v1.json 2> $OUT/bench_v1.err; echo "bench v1 rc=$?"
  for V in 1 3; do
    AGP_UMMA_V2=$V timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_v2_$V.json 2> $OUT/bench_v2_$V.err; echo "bench v2=$V rc=$?"
  done
fi
if has ns; then
  # Newton-Schulz refinement on the v2 kernel (experimental C-ABI hook) against the fp64 inverse, then the engine-integrated tail
  AGP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -x -q -s -k newton > $OUT/pytest_ns.log 2>&1; echo "ns pytest rc=$?" | tee -a $OUT/pytest_ns.log
  grep -E "rel err|mu |passed|failed|rror" $OUT/pytest_ns.log | head -20
  AGP_UMMA_V2=1 AGP_TAIL_NS=3 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_ns3.json 2> $OUT/bench_ns3.err; echo "bench ns3 rc=$?"
  tail -c 400 $OUT/bench_ns3.err
fi
if has async; then
  # host-batch steps without a per-step synchronisation (agp_step_batch_async / agp_result_wait): parity test + e2e bench leg
  AGP_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_experimental_gpu.py -m gpu -x -q -k async > $OUT/pytest_async.log 2>&1; echo "async pytest rc=$?" | tee -a $OUT/pytest_async.log
  tail -3 $OUT/pytest_async.log
  timeout 300 python bench.py --steps 100 --warmup 5 --e2e-async --no-cpu-baseline > $OUT/bench_e2e_async.json 2> $OUT/bench_e2e_async.err; echo "bench e2e-async rc=$?"
fi
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ.get("TAG", "umma_v2"), "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(os.path.basename(f), round(d["value"]), "it/s; e2e", round(d["e2e"]["value"]), "it/s;",
              {n: round(v["seconds_per_launch"] * 1e6, 1) for n, v in k.items() if "seconds_per_launch" in v})
    except Exception as e:
        print(f, "unreadable:", e)
PY
