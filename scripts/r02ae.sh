#!/bin/bash
# r02ae: warp-coherent packet traversal with shared-memory node staging (MCRT_PACKET=1: first hit only, 2: every bounce) vs per-lane traversal
TAG=r02ae
mkdir -p gpurun_out
for lib in libmcrt_pk1.so libmcrt_pk2.so; do
  echo "== $lib" | tee -a gpurun_out/${TAG}_pytest.log
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "cast_rays or full_frame or edge_sizes or degenerate or config4_small or batched or traversal_options" 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_pytest.log
done
for rep in 1 2; do
for lib in libmcrt.so libmcrt_pk1.so libmcrt_pk2.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 600 python scripts/ab_libs.py 1024 --c4 2>&1 | grep -v "^$" | tee -a gpurun_out/${TAG}_ab_packet.txt
done
done
