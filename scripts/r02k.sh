#!/bin/bash
# r02k (2 GPUs): the whole GPU suite incl. tests/test_gpu_multi.py on real NCCL ranks; post kernel A/B
TAG=r02k
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/${TAG}_gpus.txt
python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "PASSED|FAILED|SKIPPED|passed|failed" | tee gpurun_out/${TAG}_pytest_2gpu.log | tail -15
python scripts/ab_option.py post_tma=0,1 512 2>&1 | tee gpurun_out/${TAG}_ab_post_tma.txt
