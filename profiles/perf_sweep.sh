#!/bin/bash
# Device-resident A/B of the kernel build knobs (gpurun -- profiles/perf_sweep.sh [tag]): every line is one bench.py run
# (CUDA-event timing, 10 steps) of one workload with one knob changed from the default build.
#   default      parameters as kernel arguments, persistent grid-stride launch, axis-packed leaves, 4 points per thread
#   literals     every float parameter a literal in the generated source (the round-1 form)
#   nonpersist   one block per 256 groups instead of a resident grid
#   ppt1         one point per thread (quarter of the code size, no 16-byte stores)
#   noaxispack   the plain one-point leaves (no FFMA2 / FADD2)
#   unroll       octave loops fully unrolled
#   waves2       twice the resident grid
tag=${1:-r2}
out=gpurun_out/perf_sweep_${tag}.txt
mkdir -p gpurun_out
: > $out
run() { label=$1; shift; wl=$1; shift
  v=$(env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep_err.txt | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms  mm %s  sm %s MHz' % (d['value'], d['roofline']['kernel_ms'], d['run']['output_minmax'], d['clocks']['sm_mhz']))" 2>&1)
  echo "$wl $label: $v" | tee -a $out; }
for wl in ${WORKLOADS:-c4 c2a c1 c3a c5}; do
  run default $wl A=1
  run literals $wl FASTNOISE2_B200_LITERALS=1
  run nonpersist $wl FASTNOISE2_B200_PERSISTENT=0
  run ppt1 $wl FASTNOISE2_B200_PPT=1
  run noaxispack $wl FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0
  run unroll $wl 'FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_OCTAVE_LOOP=_Pragma("unroll")'
  run waves2 $wl FASTNOISE2_B200_WAVES=2
  run default_again $wl A=1
done
