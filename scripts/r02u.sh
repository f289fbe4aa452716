#!/bin/bash
# r02u (8 GPUs): the multi-GPU bench path at scale
TAG=r02u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n8.err | tee gpurun_out/${TAG}_bench_n8.json | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --contiguous 2> gpurun_out/${TAG}_n8c.err | tee gpurun_out/${TAG}_bench_n8_contiguous.json | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n4.err | tee gpurun_out/${TAG}_bench_n4.json | cut -c1-250
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${TAG}_n8.err | tail -5
