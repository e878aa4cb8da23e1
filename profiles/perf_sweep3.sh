#!/bin/bash
# Code-footprint A/B on the heavy trees (gpurun -- profiles/perf_sweep3.sh [tag]): the instruction cache decides (ncu: 70 % ICC hit
# rate, "no instruction" the largest stall for a 55 KB loop body). Variants: axis-packed or plain leaves x 4 points per thread
# unrolled (P4) / 4 points per thread in a rolled loop (P4R, a quarter of the code) / 1 point per thread (P1) x groups per thread.
tag=${1:-r2}
out=gpurun_out/perf_sweep3_${tag}.txt
mkdir -p gpurun_out
: > $out
run() { label=$1; shift; wl=$1; shift
  v=$(env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep_err.txt | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms  sm %s MHz' % (d['value'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))" 2>&1)
  echo "$wl $label: $v" | tee -a $out; }
for wl in ${WORKLOADS:-c4 c3a c5 c2b c2a c1}; do
  for it in 1 8; do
    run "pack_P4_iters$it" $wl FASTNOISE2_B200_ITERS=$it
    run "pack_P4R_iters$it" $wl FASTNOISE2_B200_ITERS=$it "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_ROLL_POINTS=1"
    run "pack_P1_iters$it" $wl FASTNOISE2_B200_ITERS=$it FASTNOISE2_B200_PPT=1
    run "plain_P4_iters$it" $wl FASTNOISE2_B200_ITERS=$it "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0"
    run "plain_P4R_iters$it" $wl FASTNOISE2_B200_ITERS=$it "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 -DFN2_ROLL_POINTS=1"
    run "plain_P1_iters$it" $wl FASTNOISE2_B200_ITERS=$it FASTNOISE2_B200_PPT=1 "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0"
  done
  run "pack_P4_iters8_barrier" $wl FASTNOISE2_B200_ITERS=8 "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_ITERATION_BARRIER=__syncthreads();"
  run "plain_P4_iters8_barrier" $wl FASTNOISE2_B200_ITERS=8 "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 -DFN2_ITERATION_BARRIER=__syncthreads();"
  run "plain_P4R_iters32" $wl FASTNOISE2_B200_ITERS=32 "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 -DFN2_ROLL_POINTS=1"
  run "plain_P4R_iters0" $wl FASTNOISE2_B200_ITERS=0 "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 -DFN2_ROLL_POINTS=1"
done
