#!/bin/bash
# A/B of library variants through bench.py (development aid): usage ab_bench.sh <frames> lib...
F=$1; shift
for lib in "$@"; do
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline --frames-per-step $F 2>&1 | tail -1 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('$lib', b['config']['frames_per_step_per_gpu'], round(b['value']), round(b['e2e']['value']), round(b['latency_mode']['frames_per_s']), b['roofline']['stage_ms'])"
done
