#!/bin/bash
# r02ap: confirm 8 resident CTAs (64 registers) for k_bounce / k_first_hit against 6 (80) and 10 (51); latency mode and 64 frames per call too
TAG=r02ap
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libmcrt.so libmcrt_c8.so libmcrt_c10.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 600 python scripts/ab_libs.py 1024 --c4 2>&1 | grep -v "^$" | tee -a gpurun_out/${TAG}_ab_bounce_regs.txt
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 600 python scripts/ab_libs.py 64 2>&1 | grep "F=64" | tee -a gpurun_out/${TAG}_ab_bounce_regs.txt
done
done
