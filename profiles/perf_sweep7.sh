#!/bin/bash
# Instruction-stream alignment A/B on the terrain tree (gpurun -- profiles/perf_sweep7.sh [tag]): block size x groups per thread x
# a block barrier at the top of every group (legal under the convergent walk: every thread of a block runs the same trip count).
tag=${1:-r2}
out=gpurun_out/perf_sweep7_${tag}.txt
mkdir -p gpurun_out
IT=FASTNOISE2_B200_ITERS; GB=FASTNOISE2_B200_GROUP_BARRIER
{
for b in 256 512 1024; do
  echo "== block $b"
  FASTNOISE2_B200_BLOCK=$b python tools/ab_sweep.py --workloads ${WORKLOADS:-c4} --variants "it1:$IT=1;it2:$IT=2;it4:$IT=4;it8:$IT=8;it0:$IT=0;bar_it2:$GB=1,$IT=2;bar_it4:$GB=1,$IT=4;bar_it8:$GB=1,$IT=8;bar_it16:$GB=1,$IT=16;bar_it0:$GB=1,$IT=0"
done
} 2>&1 | grep -v "^$" | tee $out
