#!/bin/bash
# Run on the GPU box (gpurun -- profiles/bench_lines.sh <tag>): the parity suite, then the full bench line of every workload (not
# under a profiler), the reference arm, and the pageable-destination probe. Outputs under gpurun_out/.
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputest_${tag}.txt
cat gpurun_out/gputest_${tag}.txt | tail -3
for wl in c4 c1 c2a c2b c3a c3b c5; do
  extras=""; [ $wl = c4 ] && extras="--extras"
  python bench.py --workload $wl --steps 10 --warmup 3 $extras > gpurun_out/bench_${wl}_${tag}.json 2> gpurun_out/bench_${wl}_${tag}.err
done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_c4_${tag}.json 2> gpurun_out/bench_reference_c4_${tag}.err
python bench.py > gpurun_out/bench_default_${tag}.json 2> gpurun_out/bench_default_${tag}.err
python tools/pageable_probe.py --nz 512 --calls 3 > gpurun_out/pageable_probe_${tag}.txt 2>&1
python tools/call_latency.py > gpurun_out/call_latency_${tag}.txt 2>&1
python tools/odd_width.py > gpurun_out/odd_width_${tag}.txt 2>&1
tail -c 400 gpurun_out/bench_default_${tag}.json
