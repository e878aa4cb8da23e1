#!/bin/bash
# Block size and plain-leaf A/B under the convergent walk (gpurun -- profiles/perf_sweep6.sh [tag]); one process per setting
# because both are fixed per process (block size is read once, compile flags are not part of the in-memory kernel key).
tag=${1:-r2}
out=gpurun_out/perf_sweep6_${tag}.txt
mkdir -p gpurun_out
IT=FASTNOISE2_B200_ITERS
{
for b in 64 128 256 512; do
  echo "== block $b"
  FASTNOISE2_B200_BLOCK=$b python tools/ab_sweep.py --workloads c4 --variants "it1:$IT=1;it2:$IT=2;it4:$IT=4"
  FASTNOISE2_B200_BLOCK=$b python tools/ab_sweep.py --workloads c2a,c1,c3a,c5 --variants "default:"
done
echo "== plain leaves (no axis packing), block 256"
FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 python tools/ab_sweep.py --workloads c4 --variants "it1:$IT=1;it2:$IT=2"
FASTNOISE2_B200_NVRTC_FLAGS=-DFN2_AXIS_PACK=0 python tools/ab_sweep.py --workloads c2a,c2b,c1,c5 --variants "default:"
} 2>&1 | grep -v "^$" | tee $out
