#!/bin/bash
# r02ak (8 GPUs): scaling with the direct device output
TAG=${TAG:-r02ak}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n8.err | tee gpurun_out/${TAG}_bench_n8.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n4.err | tee gpurun_out/${TAG}_bench_n4.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 5 --warmup 3 --config c3 2> gpurun_out/${TAG}_n8c3.err | tee gpurun_out/${TAG}_bench_n8_c3.json | cut -c1-200
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${TAG}_n8.err | tail -3
