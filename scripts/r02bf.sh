#!/bin/bash
# r02bf (4 GPUs): what does the exchange cost with the final code?  normal run against MCRT_DIAG_SKIP_DEPOSIT=1 (completion signal only, no RF lines moved)
TAG=r02bf
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29575 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_n4.err | tee gpurun_out/${TAG}_bench_n4.json | cut -c1-200
MCRT_DIAG_SKIP_DEPOSIT=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29576 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_n4_nodeposit.err | tee gpurun_out/${TAG}_bench_n4_nodeposit.json | cut -c1-200
