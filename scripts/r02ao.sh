#!/bin/bash
# r02ao: k_bounce register target / resident CTAs (launch bounds) with the round-2 kernel: 5 (96 regs), 6 (80, default), 7 (72), 8 (64)
TAG=r02ao
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libmcrt.so libmcrt_c5.so libmcrt_c7.so libmcrt_c8.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 600 python scripts/ab_libs.py 1024 --c4 2>&1 | grep -v "^$" | tee -a gpurun_out/${TAG}_ab_bounce_regs.txt
done
done
