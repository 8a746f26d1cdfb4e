#!/bin/bash
# per-launch speed-of-light numbers of the HBM-bound passes of one forward (run under gpurun)
TAG=${1:-r1}
timeout 500 ncu --section SpeedOfLight --clock-control none -k regex:"build_taps|stem_taps|head_conv|instnorm|l2norm" -s 120 -c 100 --csv \
    --log-file gpurun_out/sol_taps_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/sol_taps_$TAG.csv
