#!/bin/bash
# r02d: prefetch-on-push and the shared-memory short stack, each against the plain 4-wide traversal
TAG=r02d
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
for lib in libmcrt_nopf.so libmcrt.so libmcrt_ss8.so libmcrt_nopf.so libmcrt.so libmcrt_ss8.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python scripts/ab_libs.py 512 --c4 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_prefetch.txt
done
