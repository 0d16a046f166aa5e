#!/bin/bash
# round-2 evidence pass on one GPU: full GPU test suite, smoke, bench line, ncu launch list, ncu --set full of the step's top kernels and
# of the K_nm kernel at 2^20 prediction rows.   usage: gpurun --timeout 2400 -- 'bash tools/r2_evidence.sh <tag>'
TAG=${1:-ev}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -4 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 210 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 14 --warmup 3 --graph 0 --timed-only > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'tail2_step_kernel|umma_gemm_nt_kernel|umma_gemm_ps_kernel|umma_gram_tn_kernel|knm_umma_kernel|tail2_potf2_first_kernel|combine_kernel|scale_transpose_kernel|x_finalize|rowfinish' -s 60 -c 18 \
    -o $OUT/prof python bench.py --steps 6 --warmup 3 --graph 0 --timed-only > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'knm_umma_kernel' -s 4 -c 1 -o $OUT/prof_knm_1m python bench.py --predict --no-cpu-baseline > $OUT/ncu_knm.log 2>&1; echo "ncu knm rc=$?"
ls -la $OUT
