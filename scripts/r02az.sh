#!/bin/bash
# r02az: (a) accumulate: masked tail block for segment tails of >= 2 / 3 / 4 steps (-DMCRT_WIN_TAIL_BLOCK) against the default build;
#        (b) e2e: 1 / 2 / 4 calls per step in the streaming driver
TAG=r02az
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libmcrt.so libmcrt_tb2.so libmcrt_tb3.so libmcrt_tb4.so; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib timeout 300 python scripts/ab_libs.py 1024 2>&1 | grep "F=" | tee -a gpurun_out/${TAG}_ab.txt
done
done
MCRT_LIB_PATH=$PWD/mcray_tracing_b200/libmcrt_tb2.so python -m pytest tests -m gpu -x -q -k "accumulate or full_frame or edge_sizes or ray_tree or configs or streaming" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_tb2.log
for sb in 1 2 4; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-sub-batches $sb 2>> gpurun_out/${TAG}_bench.err | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().split('\n')[-1]); e = b['e2e']
print('e2e sub_batches=$sb', round(e['value'], 1), 'frames/s  frac of PCIe ceiling', round(e['roofline']['frac'], 3), ' device value', round(b['value'], 1))" | tee -a gpurun_out/${TAG}_e2e.txt
done
