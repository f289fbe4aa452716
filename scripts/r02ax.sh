#!/bin/bash
# r02ax: (a) accumulate window-edge skip (default build) against the previous kernel (libmcrt_noskip.so);
#        (b) traversal: 64-bit stack entries with pop-time culling (cull), interval slack folded into the far-plane constants (fold), both
TAG=r02ax
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libmcrt.so libmcrt_noskip.so libmcrt_cull.so libmcrt_fold.so libmcrt_both.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 300 python scripts/ab_libs.py 1024 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
done
done
for lib in libmcrt.so libmcrt_both.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 300 python scripts/ab_libs.py 8 --c4 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_default.log
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt_both.so python -m pytest tests -m gpu -x -q -k "closest_hit or cast_rays or traversal or ray_tree or full_frame" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_both.log
