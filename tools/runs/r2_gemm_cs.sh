#!/bin/bash
# ring-depth variants of the first-generation GEMM kernel (Gram product): standalone harness
OUT=gpurun_out/${1:-gemm_cs}; mkdir -p $OUT
for v in rs4cs2 rs4cs3 rs3cs3; do
  echo "=== $v (pre-split kernel for the triangular products)"; timeout 60 ./profiles/microbench/gemm_bench_$v 8192 512 2>&1 | tee $OUT/gemm_bench_$v.txt
  echo "=== $v AGP_UMMA_PS=0"; AGP_UMMA_PS=0 timeout 60 ./profiles/microbench/gemm_bench_$v 8192 512 2>&1 | tee $OUT/gemm_bench_${v}_ps0.txt
done
echo "=== C3 rs4cs3"; timeout 60 ./profiles/microbench/gemm_bench_rs4cs3 16384 1024 2>&1 | tee $OUT/gemm_bench_rs4cs3_c3.txt
