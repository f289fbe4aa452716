#!/bin/bash
TAG=r02m
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_c5_counters.csv \
    python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c5.log 2>&1
