#!/bin/bash
# r02ay: (a) background tree optimisation (default) against the plain device LBVH (bvh_optimise=0); (b) accumulate window-edge skip in the
# checked-step loop (default build) against the previous kernel (libmcrt_noskip.so); GPU suite; bench line
TAG=r02ay
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt.so timeout 300 python scripts/ab_libs.py 1024 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt.so timeout 300 python scripts/ab_libs.py 1024 --opt bvh_optimise=0 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt_noskip.so timeout 300 python scripts/ab_libs.py 1024 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
done
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt.so timeout 300 python scripts/ab_libs.py 8 --c4 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt.so timeout 300 python scripts/ab_libs.py 8 --c4 --opt bvh_optimise=0 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-400
