#!/bin/bash
# ncu captures behind the numbers in profiles/ (run under gpurun, 1 GPU).  Usage: bash tools/profile.sh <tag>
# Numbers printed by a run under ncu are never bench values.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
FWD="python tools/one_forward.py --forwards 3"
# 1. every launch of three forwards with its device time (cold-cache, serialised: compare SHARES); the summariser keeps
#    the last forward (weights are packed during the first)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_$TAG.csv $FWD > /dev/null 2>&1
# 2. the dominant kernel: conv_gemm<256> on a 512->512 3x3 layer of img_enc (96 samples), steady state
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel.*256" -s 45 -c 1 \
    -f -o $OUT/prof_conv_$TAG $FWD > /dev/null 2>&1
# 3. the correlation kernels: tensor-core tiles, finish (merge + gather), operand normalisation
timeout 400 ncu --set full --clock-control none --import-source on -k regex:corr_tile -s 2 -c 1 \
    -f -o $OUT/prof_corr_$TAG $FWD > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:warp_mean_taps -s 2 -c 1 \
    -f -o $OUT/prof_finish_$TAG $FWD > /dev/null 2>&1
# 4. the vertical-reuse stem kernel (96 samples: the first of the two launches per forward)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_vr -s 4 -c 1 \
    -f -o $OUT/prof_stem_$TAG $FWD > /dev/null 2>&1
ls -la $OUT
