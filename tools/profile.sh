#!/bin/bash
# ncu captures behind the numbers in profiles/ (run under gpurun, 1 GPU).  Usage: bash tools/profile.sh <tag>
# Numbers printed by a run under ncu are never bench values.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# 1. every launch of one steady-state forward with its device time (cold-cache, serialised: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 140 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH > /dev/null 2>&1
# 2. the dominant kernel: conv_gemm<256> on the 512->512 3x3 layer of img_enc (96 samples)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 50 -c 1 \
    -f -o $OUT/prof_conv_$TAG $BENCH > /dev/null 2>&1
# 3. the fused correlation + warp kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:corr_warp -s 1 -c 1 \
    -f -o $OUT/prof_corr_$TAG $BENCH > /dev/null 2>&1
ls -la $OUT
