#!/bin/bash
# r02ad (2 GPUs): the whole GPU suite incl. tests/test_gpu_multi.py, and the multi-GPU bench with the final code
TAG=r02ad
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2.json | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>> gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2_reference.json | cut -c1-200
