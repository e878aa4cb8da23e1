#!/bin/bash
# Second code-shape A/B (gpurun -- profiles/perf_sweep5.sh [tag]): the convergent group walk with unconditional min/max against
# the divergent walk on every workload, and groups per thread for the terrain tree under the convergent walk.
tag=${1:-r2}
out=gpurun_out/perf_sweep5_${tag}.txt
mkdir -p gpurun_out
CV=FASTNOISE2_B200_CONVERGENT; IT=FASTNOISE2_B200_ITERS
{
python tools/ab_sweep.py --workloads c4 --variants "div_it1:$CV=0,$IT=1;conv_it1:$IT=1;conv_it2:$IT=2;conv_it3:$IT=3;conv_it4:$IT=4;conv_it8:$IT=8;conv_default:;div_it1_again:$CV=0,$IT=1;conv_it1_again:$IT=1"
python tools/ab_sweep.py --workloads c2a,c2b,c1,c3a,c3b,c5 --variants "div:$CV=0;conv:;conv_it4:$IT=4;conv_it16:$IT=16;div_again:$CV=0;conv_again:"
} 2>&1 | tee $out
