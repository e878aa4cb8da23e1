#!/bin/bash
# scratch: throughput under different codegen knobs
run() { # label, env..., workload
  label=$1; shift; wl=$1; shift
  v=$(env "$@" python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms' % (d['value'], d['roofline']['kernel_ms']))")
  echo "$wl $label: $v"
}
for wl in c1 c2a c3a c4 c5; do
  for p in 4 2 1; do run "ppt=$p" $wl FASTNOISE2_B200_PPT=$p; done
done
for wl in c2a c4; do
  run "octaves unrolled" $wl "FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_OCTAVE_LOOP="
done
for wl in c1 c3a c4; do
  run "relaxed math" $wl FN2_DUMMY=1 
done
