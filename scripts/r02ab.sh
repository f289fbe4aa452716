#!/bin/bash
# r02ab: history-grouped compaction (refracted survivors before reflected ones) A/B
TAG=r02ab
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "traversal_options or cast_rays or full_frame or batched or edge_sizes" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
python scripts/ab_option.py group_histories=0,1 1024 2>&1 | tee gpurun_out/${TAG}_ab_group_histories.txt
python scripts/ab_option.py group_histories=0,1 64 2>&1 | tee -a gpurun_out/${TAG}_ab_group_histories.txt
for v in 0 1; do python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --option group_histories=$v 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read()); print('c4 group_histories=$v', round(b['value']), {k: round(x, 3) for k, x in b['roofline']['stage_ms'].items()})" | tee -a gpurun_out/${TAG}_ab_group_histories.txt; done
