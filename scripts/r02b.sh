#!/bin/bash
# r02b: barrier-free ordered compaction + OR-formed node addresses + FMA minimax numerics, against the round-1 library
TAG=r02b
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
for lib in libmcrt_r01.so libmcrt.so; do
  echo "== $lib" | tee -a gpurun_out/${TAG}_ab.txt
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python scripts/ab_option.py ordered_compaction=1,0 512 2>&1 | tee -a gpurun_out/${TAG}_ab.txt
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
