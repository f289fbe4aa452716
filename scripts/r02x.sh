#!/bin/bash
# r02x: envelope tile kernel variants (chunks per warp, register target) on config 5
TAG=r02x
mkdir -p gpurun_out
for lib in libmcrt.so libmcrt_b3.so libmcrt_c16.so libmcrt_c16b3.so; do
  for rep in 1 2; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}.err | python -c "
import sys, json
b = json.loads(sys.stdin.read())
r = b['roofline']
print('$lib', 'ms/step', round(b['ms_per_step'], 3), 'post', round(r['stage_ms']['post'], 4), 'frac', round(r['frac_of_max_bytes_flops_roof'], 3))
" | tee -a gpurun_out/${TAG}_ab_envelope_tiles.txt
  done
done
tail -2 gpurun_out/${TAG}.err
