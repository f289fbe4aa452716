#!/bin/bash
# r02a: round-2 baseline on today's box -- GPU tests, bench line, one full-set ncu capture WITH source of every
# kernel of one step (k_bounce x10, k_accumulate_win, k_post_fused) for the instruction attribution.
TAG=r02a
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
ncu --set full --clock-control none --import-source on -k regex:'k_bounce|k_accumulate_win|k_post_fused|k_first_hit' -s 0 -c 13 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
