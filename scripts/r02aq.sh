#!/bin/bash
TAG=r02aq
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest.log
(python scripts/ab_option.py tail_merge=1,2 1024; python scripts/ab_option.py tail_merge=1,2 64; python scripts/ab_option.py tail_merge=1,2 1) 2>&1 | tee gpurun_out/${TAG}_ab_tail.txt
