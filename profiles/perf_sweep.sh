#!/bin/bash
run() { label=$1; shift; wl=$1; shift
  v=$(python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f Mpts/s  kernel %.3f ms mm %s' % (d['value'], d['roofline']['kernel_ms'], d['config']['output_minmax']))")
  echo "$wl $label: $v"; }
for wl in c1 c2a c2b c3a c4 c5; do
  run "strict generic" $wl --math strict_generic
  run "strict tables " $wl --math strict
  run "relaxed tables" $wl --math relaxed
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
