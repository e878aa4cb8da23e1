#!/bin/bash
# Launch-geometry and register-cap A/B (gpurun -- profiles/perf_sweep2.sh [tag]); see perf_sweep.sh for the format.
#   itersN       every thread walks about N groups of its call (0 = exactly the resident grid, 1 = one group per thread)
#   minblocksN   __launch_bounds__( 256, N ): register cap 65536 / ( 256 N )
tag=${1:-r2}
out=gpurun_out/perf_sweep2_${tag}.txt
mkdir -p gpurun_out
: > $out
run() { label=$1; shift; wl=$1; shift
  v=$(env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep_err.txt | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms  mm %s  sm %s MHz' % (d['value'], d['roofline']['kernel_ms'], d['run']['output_minmax'], d['clocks']['sm_mhz']))" 2>&1)
  echo "$wl $label: $v" | tee -a $out; }
for wl in ${WORKLOADS:-c4 c3a c2a c2b c1 c5}; do
  for it in 0 1 2 4 8 16 64; do run iters$it $wl FASTNOISE2_B200_ITERS=$it; done
  for mb in 2 3 5 6; do run minblocks$mb $wl FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_MIN_BLOCKS_ARG=,$mb; done
  run ppt1 $wl FASTNOISE2_B200_PPT=1
  run noaxispack $wl FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0
done
