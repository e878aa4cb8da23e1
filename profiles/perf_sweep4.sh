#!/bin/bash
# Code-shape A/B (gpurun -- profiles/perf_sweep4.sh [tag]): one-body octave loops x convergent group walk x groups per thread, all
# workloads, one process per workload (tools/ab_sweep.py; the cubins of every variant are compiled ahead into the in-tree cache).
tag=${1:-r2}
out=gpurun_out/perf_sweep4_${tag}.txt
mkdir -p gpurun_out
OB=FASTNOISE2_B200_ONE_BODY; CV=FASTNOISE2_B200_CONVERGENT; IT=FASTNOISE2_B200_ITERS
{
python tools/ab_sweep.py --workloads c4 --variants "base_it1:$OB=0,$CV=0,$IT=1;conv_it1:$OB=0,$CV=1,$IT=1;one_it1:$OB=1,$CV=0,$IT=1;both_it1:$OB=1,$CV=1,$IT=1;both_it2:$IT=2;both_it4:$IT=4;both_it8:$IT=8;both_it0:$IT=0;one_it8:$OB=1,$CV=0,$IT=8;both_plain_it1:$IT=1,FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0;both_plain_it8:$IT=8,FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0;base_it1_again:$OB=0,$CV=0,$IT=1;both_it1_again:$IT=1"
python tools/ab_sweep.py --workloads c2a,c2b,c1,c3a,c3b,c5 --variants "base:$OB=0,$CV=0;conv:$OB=0,$CV=1;one:$OB=1,$CV=0;both:;both_it1:$IT=1;both_it2:$IT=2;both_it4:$IT=4;both_plain:FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0;base_again:$OB=0,$CV=0"
} 2>&1 | tee $out
