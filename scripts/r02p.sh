#!/bin/bash
TAG=r02p
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --contiguous 2> gpurun_out/${TAG}_n2c.err | tee gpurun_out/${TAG}_bench_n2_contiguous.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --config c3 2> gpurun_out/${TAG}_n2c3.err | tee gpurun_out/${TAG}_bench_n2_c3.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/${TAG}_n2r.err | cut -c1-200
tail -n 5 gpurun_out/${TAG}_n2.err gpurun_out/${TAG}_n2c.err gpurun_out/${TAG}_n2c3.err | grep -v "^$" | tail -20
