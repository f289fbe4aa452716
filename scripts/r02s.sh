#!/bin/bash
TAG=r02s
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python scripts/ab_option.py post_tma=1,1 512 2>&1 | tee gpurun_out/${TAG}_stage.txt
