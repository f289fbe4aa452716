#!/bin/bash
TAG=r02af
mkdir -p gpurun_out
for f in 8 16 32; do
python bench.py --config c5 --frames-per-step $f --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read()); r = b['roofline']
print('c5 F=$f', round(b['value'], 1), 'frames/s', round(b['ms_per_step'], 3), 'ms/step', {k: round(x, 3) for k, x in r['stage_ms'].items()}, 'e2e', round(b['e2e']['value'], 1), 'post frac', round(r['frac_of_max_bytes_flops_roof'], 3), [ (s['stage'], round(s.get('issue_frac') or 0, 3)) for s in r['stages']])" | tee -a gpurun_out/${TAG}_c5_frames_per_step.txt
done
for f in 64 128; do
python bench.py --config c4 --frames-per-step $f --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read()); r = b['roofline']
print('c4 F=$f', round(b['value'], 1), 'frames/s', round(b['ms_per_step'], 3), 'ms/step', {k: round(x, 3) for k, x in r['stage_ms'].items()}, 'e2e', round(b['e2e']['value'], 1))" | tee -a gpurun_out/${TAG}_c5_frames_per_step.txt
done
