#!/bin/bash
# Scalar gradients under packed consumers (gpurun -- profiles/perf_sweep8.sh [tag]): -DFN2_PAIR_TABLE_LOADS=0 against the default, twice.
tag=${1:-r2}
IT=FASTNOISE2_B200_ITERS
for f in "" "-DFN2_PAIR_TABLE_LOADS=0" "" "-DFN2_PAIR_TABLE_LOADS=0"; do
  echo "== flags [$f]"
  FASTNOISE2_B200_NVRTC_FLAGS="$f" python tools/ab_sweep.py --workloads c4,c1 --variants "it1:$IT=1;it2:$IT=2;default:" 2>&1 | grep -v "^$"
done | tee gpurun_out/perf_sweep8_${tag}.txt
