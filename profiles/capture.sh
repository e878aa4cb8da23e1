#!/bin/bash
# Run on the GPU box (gpurun -- profiles/capture.sh <tag>): launch list of the bench command, one full ncu capture of
# the dominant kernel per workload, and untimed-by-ncu bench lines for every workload. Outputs under gpurun_out/.
tag=${1:-r1}
mkdir -p gpurun_out
for wl in c1 c2a c2b c3a c3b c4 c5; do
  python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${wl}_${tag}.json 2> gpurun_out/bench_${wl}_${tag}.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_${tag}.csv \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c4_${tag}.log 2>&1
for wl in c4 c2a c5 c3a c1; do
  ncu --set full --clock-control none --import-source on -k fn2_entry -c 1 -f -o gpurun_out/prof_${wl}_${tag} \
      python bench.py --workload $wl --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${wl}_${tag}.log 2>&1
done
ls -la gpurun_out
