#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8 -- profiles/scaling.sh <tag>): the bench line at N = 8 and 4 (strong scaling headline, weak
# next to it, optional gather: NCCL and fused), the one-process fan-out at 8 visible GPUs, the bare D2H ceiling, the multi-GPU tests.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${tag}.txt 2>&1
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --extras \
      > gpurun_out/scale_n${n}_${tag}.json 2> gpurun_out/scale_n${n}_${tag}.err
done
python bench.py --gpus 1 --steps 5 --extras --no-cpu-baseline > gpurun_out/scale_n1_fanout_${tag}.json 2> gpurun_out/scale_n1_fanout_${tag}.err
python tools/d2h_bandwidth.py > gpurun_out/d2h_bandwidth_8gpu_${tag}.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -q -rf 2>&1 | tail -15 > gpurun_out/gputest_multi8_${tag}.txt
tail -c 600 gpurun_out/scale_n8_${tag}.json; cat gpurun_out/gputest_multi8_${tag}.txt | tail -3; cat gpurun_out/d2h_bandwidth_8gpu_${tag}.txt
