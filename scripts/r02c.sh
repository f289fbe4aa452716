#!/bin/bash
# r02c: 8-wide BVH with octant-ordered group stack (default build) against the sorted 4-wide traversal (-DMCRT_BVH8=0)
TAG=r02c
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
for lib in libmcrt_bvh4.so libmcrt.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python scripts/ab_libs.py 512 --c4 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_bvh8.txt
done
