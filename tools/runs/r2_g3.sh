#!/bin/bash
export TAG=${1:-g3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "hyperparameter_training or hundred_iterations or online" > $OUT/pytest_sel.log 2>&1; echo "selected pytest rc=$?"; tail -6 $OUT/pytest_sel.log
timeout 120 ./profiles/microbench/gemm_trace > $OUT/gemm_trace.txt 2>&1; echo "trace rc=$?"; tail -c 300 $OUT/gemm_trace.txt
timeout 600 python bench.py --predict > $OUT/bench_predict_1m.json 2> $OUT/bench_predict_1m.err; echo "predict rc=$?"; tail -c 600 $OUT/bench_predict_1m.err
timeout 600 python bench.py --predict --predict-chunk 8192 > $OUT/bench_predict_8k.json 2> $OUT/bench_predict_8k.err; echo "predict 8k rc=$?"; tail -c 600 $OUT/bench_predict_8k.err
head -c 1500 $OUT/bench_predict_1m.json; echo; head -c 600 $OUT/bench_predict_8k.json
