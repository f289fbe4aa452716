#!/bin/bash
# r02bc (8 GPUs): final code, weak-scaling bench line at N = 8 (bit-identity check inside)
TAG=r02bc
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n8.err | tee gpurun_out/${TAG}_bench_n8.json | cut -c1-250
