#!/bin/bash
# r02z: after removing the coherence-sort branch from k_bounce: tests + timing against r02v (111.6 k frames/s)
TAG=r02z
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', round(b['value']), 'ms/step', round(b['ms_per_step'], 3), {k: round(v, 3) for k, v in b['roofline']['stage_ms'].items()}, 'e2e', round(b['e2e']['value']))
"
done
