#!/bin/bash
TAG=r02at
mkdir -p gpurun_out
(python scripts/ab_option.py tail_merge=1,12,20 512) 2>&1 | tee gpurun_out/${TAG}_ab_tail.txt
for c in c4 c5 c3; do for v in 1 12; do
python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --option tail_merge=$v 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read()); print('$c tail_merge=$v', round(b['value'], 1), 'frames/s', {k: round(x, 3) for k, x in b['roofline']['stage_ms'].items()})" | tee -a gpurun_out/${TAG}_ab_tail.txt
done; done
for v in 1 12; do python bench.py --config c4 --frames-per-step 64 --steps 5 --warmup 3 --no-cpu-baseline --option tail_merge=$v 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read()); print('c4 F=64 tail_merge=$v', round(b['value'], 1), 'frames/s', {k: round(x, 3) for k, x in b['roofline']['stage_ms'].items()})" | tee -a gpurun_out/${TAG}_ab_tail.txt
done
