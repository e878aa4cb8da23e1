#!/bin/bash
# Run on the GPU box (gpurun -- profiles/recapture_c2b_c5.sh <tag>): the parity suite, then bench line + ncu capture of the two
# workloads whose kernels changed last (tileable and position-array entries: __grid_constant__ argument structs), and the no-flag line.
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputest_${tag}.txt
tail -3 gpurun_out/gputest_${tag}.txt
for wl in c2b c5; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_${tag}.json 2> gpurun_out/bench_${wl}_${tag}.err
  ncu --set full --clock-control none --import-source on -k fn2_entry -c 1 -f -o gpurun_out/prof_${wl}_${tag} \
      python bench.py --workload $wl --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${wl}_${tag}.log 2>&1
done
python bench.py > gpurun_out/bench_default_${tag}.json 2> gpurun_out/bench_default_${tag}.err
python tools/ab_sweep.py --workloads c2b,c5,c4,c1 --variants "final:" 2>&1 | grep -v "^$" > gpurun_out/final_check_${tag}.txt
cat gpurun_out/final_check_${tag}.txt; tail -c 300 gpurun_out/bench_default_${tag}.json
