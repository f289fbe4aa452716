#!/bin/bash
# r02l: the new bench.py (all configs) + stage counters
TAG=r02l
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-200
for c in c3 c4 c5; do
  python bench.py --config $c --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_$c.err | tee gpurun_out/${TAG}_bench_$c.json | cut -c1-300
done
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_counters.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_counters.log 2>&1
tail -3 gpurun_out/*.err
