#!/bin/bash
# r02g: TMA-staged fused post kernel (k_post_tma) -- parity, A/B against round 1's k_post_fused (option post_tma), bench
TAG=r02i
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.log
python scripts/ab_option.py post_tma=0,1 512 2>&1 | tee gpurun_out/${TAG}_ab_post_tma.txt
python scripts/ab_option.py post_tma=0,1 1 2>&1 | tee -a gpurun_out/${TAG}_ab_post_tma.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-400
