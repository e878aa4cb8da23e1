#!/bin/bash
# Device-resident throughput of every workload in the three math modes (gpurun -- profiles/perf_sweep.sh).
run() { label=$1; shift; wl=$1; shift
  v=$(python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms mm %s' % (d['value'], d['roofline']['kernel_ms'], d['config']['output_minmax']))")
  echo "$wl $label: $v"; }
for wl in ${WORKLOADS:-c1 c2a c2b c3a c3b c4 c5}; do
  for mode in ${MODES:-strict_generic strict relaxed}; do run "$mode" $wl --math $mode; done
done
