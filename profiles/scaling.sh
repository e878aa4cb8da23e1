#!/bin/bash
# Weak-scaling bench on one 8-GPU box (gpurun --gpus 8 -- profiles/scaling.sh <tag>): N = 1, 2, 4, 8 back to back.
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1_${tag}.json 2> gpurun_out/scale_n1_${tag}.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n${n}_${tag}.json 2> gpurun_out/scale_n${n}_${tag}.err
done
nvidia-smi topo -m > gpurun_out/topo_${tag}.txt 2>&1
for n in 1 2 4 8; do tail -1 gpurun_out/scale_n${n}_${tag}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, 'ms/step %.2f' % d['ms_per_step'])"; done
