#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line, ncu launch list, ncu --set full of the top kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -4 $OUT/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
# launch list: 10 whole steps of the timed loop (21 launches per step; skip upload/refresh_K/warm-up launches)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 210 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 14 --warmup 3 --graph 0 --timed-only > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'tail2_step_kernel|umma_gemm_nt_kernel|knm_umma_kernel|tail2_potf2_first_kernel|combine_kernel|scale_transpose_kernel' -s 60 -c 15 \
    -o $OUT/prof python bench.py --steps 6 --warmup 3 --graph 0 --timed-only > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
