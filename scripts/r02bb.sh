#!/bin/bash
# r02bb (2 GPUs): final code -- multi-GPU test + N=2 bench line (bit-identity check inside)
TAG=r02bb
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2.json | cut -c1-250
