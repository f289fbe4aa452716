#!/bin/bash
# r02ah (2 GPUs): what does the gather cost per step?  default deposit vs NCCL gather vs no exchange at all (diagnostic)
TAG=r02ah
mkdir -p gpurun_out
for g in p2p none nccl p2p none; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --gather $g 2> gpurun_out/${TAG}_$g.err | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gather=$g', round(b['value']), 'ms/step', round(b['ms_per_step'], 3), [round(x, 3) for x in b['ms_per_step_per_rank']], b.get('n_gpu_bit_identical'))" | tee -a gpurun_out/${TAG}_gather_cost.txt
done
tail -3 gpurun_out/${TAG}_none.err
