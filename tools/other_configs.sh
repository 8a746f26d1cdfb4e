#!/bin/bash
# BASELINE.json configs 3 (pose) and 5 (n_source sweep at bs=16) + the bf16x3 point (run under gpurun, 1 GPU)
OUT=gpurun_out; mkdir -p $OUT
F="--steps 10 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline --no-fast-point"
timeout 300 python bench.py $F --pose > $OUT/oc_pose.json 2> /dev/null
for n in 1 3 5 8; do timeout 300 python bench.py $F --batch 16 --n-source $n > $OUT/oc_n$n.json 2> /dev/null; done
timeout 300 python bench.py $F --math bf16x3 > $OUT/oc_bf16x3.json 2> /dev/null
for f in $OUT/oc_*.json; do echo $f; cut -c1-160 $f; done
