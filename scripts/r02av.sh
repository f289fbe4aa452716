#!/bin/bash
# r02av: accumulate kernel launch bounds / unroll revisited with the round-2 kernel
TAG=r02av
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libmcrt.so libmcrt_a6.so libmcrt_a7.so libmcrt_a10.so libmcrt_u6.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 600 python scripts/ab_libs.py 1024 2>&1 | grep "F=1024" | tee -a gpurun_out/${TAG}_ab_accumulate_regs.txt
done
done
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt_u6.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulate or full_frame or edge_sizes" 2>&1 | tail -2
