#!/bin/bash
# Run on the GPU box (gpurun -- profiles/capture.sh <tag>): the parity suite; for every workload the full bench line (not under ncu)
# and one `ncu --set full` capture of the dominant kernel; the launch list of the bench command; the probes and harnesses.
# Outputs under gpurun_out/; profiles/summarize_ncu.py + attribute_sass.py turn the reports into the tracked summaries.
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputest_${tag}.txt
tail -3 gpurun_out/gputest_${tag}.txt
python tools/alu_probe.py > gpurun_out/alu_probe_${tag}.txt 2>&1
for wl in c4 c1 c2a c2b c3a c3b c5; do
  extras=""; [ $wl = c4 ] && extras="--extras"
  python bench.py --workload $wl --steps 10 --warmup 3 $extras > gpurun_out/bench_${wl}_${tag}.json 2> gpurun_out/bench_${wl}_${tag}.err
done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_c4_${tag}.json 2> gpurun_out/bench_reference_c4_${tag}.err
python bench.py > gpurun_out/bench_default_${tag}.json 2> gpurun_out/bench_default_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_${tag}.csv \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c4_${tag}.log 2>&1
for wl in c4 c1 c2a c2b c3a c3b c5; do
  ncu --set full --clock-control none --import-source on -k fn2_entry -c 1 -f -o gpurun_out/prof_${wl}_${tag} \
      python bench.py --workload $wl --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${wl}_${tag}.log 2>&1
done
python tools/call_latency.py > gpurun_out/call_latency_${tag}.txt 2>&1
python tools/d2h_bandwidth.py > gpurun_out/d2h_bandwidth_${tag}.txt 2>&1
python tools/pageable_probe.py --nz 512 --calls 3 > gpurun_out/pageable_probe_${tag}.txt 2>&1
python tools/odd_width.py > gpurun_out/odd_width_${tag}.txt 2>&1
python tools/node_benchmark.py --filter "node trees" > gpurun_out/node_benchmark_trees_${tag}.txt 2>&1
python tools/node_benchmark.py --iterations 10 > gpurun_out/node_benchmark_${tag}.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.txt 2>&1
tail -c 300 gpurun_out/bench_default_${tag}.json; cat gpurun_out/smoke_${tag}.txt
