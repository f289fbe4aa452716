#!/bin/bash
# r02bd: accumulate -- the lane's ring-column address pinned in a register (-DMCRT_WIN_PIN_COL=1: 265 instead of 273 instructions on the block's fast path)
TAG=r02bd
mkdir -p gpurun_out
for rep in 1 2 3; do
for lib in libmcrt.so libmcrt_pin.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 300 python scripts/ab_libs.py 1024 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
done
done
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt_pin.so python -m pytest tests -m gpu -x -q -k "accumulate or full_frame or edge_sizes or ray_tree" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_pin.log
