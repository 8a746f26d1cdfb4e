#!/bin/bash
# Round-2 GPU session script (run under gpurun, 1 GPU):  bash tools/gpu_r2.sh <tag> [stages...]
# stages: tests bench ab launches ncu_wino   (default: all).  Numbers printed under ncu are never bench values.
TAG=${1:-r2}; shift
STAGES=${@:-tests bench ab launches ncu_wino}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
FWD="python tools/one_forward.py --forwards 3"
for st in $STAGES; do
  case $st in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/pytest_$TAG.log 2>&1
      echo "pytest exit $?" >> $OUT/pytest_$TAG.log; tail -15 $OUT/pytest_$TAG.log ;;
    tests_all)
      timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_$TAG.log 2>&1
      echo "pytest exit $?" >> $OUT/pytest_$TAG.log; tail -25 $OUT/pytest_$TAG.log ;;
    bench)
      timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
      tail -3 $OUT/bench_$TAG.err; cut -c1-600 $OUT/bench_$TAG.json ;;
    ab)
      timeout 400 python bench.py --steps 10 --warmup 3 --no-winograd --no-cpu-baseline > $OUT/bench_nowino_$TAG.json 2> $OUT/bench_nowino_$TAG.err
      cut -c1-300 $OUT/bench_nowino_$TAG.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
          --log-file $OUT/launches_$TAG.csv $FWD > /dev/null 2>&1 ;;
    ncu_wino)
      # Winograd GEMM (2-CTA kernel; the first conv_gemm2 launches of a forward are the 64->128/128->256 stride-2 convs:
      # skip into the ResnetBlock region), input and output transform passes
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm2_kernel -s 45 -c 1 \
          -f -o $OUT/prof_winogemm_$TAG $FWD > /dev/null 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:wino_input_kernel -s 30 -c 1 \
          -f -o $OUT/prof_winoin_$TAG $FWD > /dev/null 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:wino_output_kernel -s 30 -c 1 \
          -f -o $OUT/prof_winoout_$TAG $FWD > /dev/null 2>&1 ;;
    tests_new)
      timeout 900 python -m pytest tests -m gpu -q --durations=5 -k "postprocess or uint8 or wino or invalidate or bs32 or direct_and or golden or two_cta or stem or classmap" > $OUT/pytest_new_$TAG.log 2>&1
      echo "pytest exit $?" >> $OUT/pytest_new_$TAG.log; tail -25 $OUT/pytest_new_$TAG.log ;;
    parity)
      timeout 600 python tools/parity_report.py --winograd bridge --chunk-kb 2 4 8 > $OUT/parity_$TAG.log 2>&1
      timeout 300 python tools/parity_report.py --winograd off >> $OUT/parity_$TAG.log 2>&1
      grep -E "WORST|Error|error" $OUT/parity_$TAG.log ;;
    parity_mix)
      timeout 700 python tools/parity_report.py --winograd bridge --chunk-kb img_enc=2,default=4 > $OUT/parity_$TAG.log 2>&1
      grep -E "WORST|Error|error" $OUT/parity_$TAG.log ;;
    parity_sf)
      timeout 900 python tools/parity_report.py --winograd bridge --small-first --chunk-kb img_enc=2,default=4 img_enc=3,default=4 4 > $OUT/parity_$TAG.log 2>&1
      grep -E "WORST|Error|error" $OUT/parity_$TAG.log ;;
    bench_nostem)
      timeout 400 python bench.py --steps 10 --warmup 3 --direct-stem --no-cpu-baseline --no-torch-cuda-baseline --no-fast-point > $OUT/bench_nostem_$TAG.json 2> $OUT/bench_nostem_$TAG.err
      cut -c1-300 $OUT/bench_nostem_$TAG.json ;;
    tests_corr)
      timeout 900 python -m pytest tests -m gpu -q --durations=5 -k "corr or stem or bridge or golden or train_mode or cache or full_batch" > $OUT/pytest_corr_$TAG.log 2>&1
      echo "pytest exit $?" >> $OUT/pytest_corr_$TAG.log; tail -25 $OUT/pytest_corr_$TAG.log ;;
    demo)
      timeout 300 python tools/demo_point.py > $OUT/demo_$TAG.log 2>&1; tail -6 $OUT/demo_$TAG.log ;;
    bench_c3)
      timeout 400 python - > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err <<'PY'
import sys, runpy
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-torch-cuda-baseline", "--no-fast-point"]
import wacv23_tsnet_b200.engine as E
_orig = E.ForwardEngine.__init__
def _init(self, *a, **k):
    _orig(self, *a, **k)
    self.wino_chunk_kb = {"img_enc": 3, "default": 4}
E.ForwardEngine.__init__ = _init
runpy.run_path("bench.py", run_name="__main__")
PY
      cut -c1-300 $OUT/bench_c3_$TAG.json ;;
    bench_bv1)
      timeout 400 python bench.py --steps 10 --warmup 3 --bridge-variant 1 --no-cpu-baseline --no-torch-cuda-baseline --no-fast-point > $OUT/bench_bv1_$TAG.json 2> $OUT/bench_bv1_$TAG.err
      cut -c1-300 $OUT/bench_bv1_$TAG.json ;;
    bench_c4)
      timeout 400 python bench.py --steps 10 --warmup 3 --wino-chunk-kb 4 --no-cpu-baseline --no-torch-cuda-baseline --no-fast-point > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err
      cut -c1-300 $OUT/bench_c4_$TAG.json ;;
    tests_wino)
      timeout 900 python -m pytest tests -m gpu -q --durations=5 -k "wino or direct_and or golden or cuda_graph or train_mode or full_batch" > $OUT/pytest_wino_$TAG.log 2>&1
      echo "pytest exit $?" >> $OUT/pytest_wino_$TAG.log; tail -25 $OUT/pytest_wino_$TAG.log ;;
    bench_unfused)
      timeout 400 python bench.py --steps 10 --warmup 3 --winograd-unfused --no-cpu-baseline --no-torch-cuda-baseline --no-fast-point > $OUT/bench_unfused_$TAG.json 2> $OUT/bench_unfused_$TAG.err
      cut -c1-300 $OUT/bench_unfused_$TAG.json ;;
    ncu_bridge)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:wino_bridge_kernel -s 30 -c 2 \
          -f -o $OUT/prof_bridge_$TAG $FWD > /dev/null 2>&1 ;;
    winobench)
      timeout 300 python tools/wino_bench.py > $OUT/winobench_$TAG.log 2>&1; cat $OUT/winobench_$TAG.log ;;
    ncu_stem)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_vr_kernel -s 4 -c 1 \
          -f -o $OUT/prof_stem_$TAG $FWD > /dev/null 2>&1 ;;
    ncu_corr)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:"corr_|l2norm|warp_mean" -s 6 -c 6 \
          -f -o $OUT/prof_corr_$TAG $FWD > /dev/null 2>&1 ;;
  esac
done
ls -la $OUT | tail -20
