#!/bin/bash
TAG=r02t
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_accumulate_win' -s 2 -c 1 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
