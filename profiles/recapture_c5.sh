tag=r2q
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_${tag}.json 2> gpurun_out/bench_c5_${tag}.err
ncu --set full --clock-control none --import-source on -k fn2_entry -c 1 -f -o gpurun_out/prof_c5_${tag} python bench.py --workload c5 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5_${tag}.log 2>&1
python tools/node_benchmark.py --filter "node trees" > gpurun_out/node_benchmark_trees_${tag}.txt 2>&1
python tools/node_benchmark.py --iterations 10 > gpurun_out/node_benchmark_${tag}.txt 2>&1
tail -c 300 gpurun_out/bench_c5_${tag}.json
