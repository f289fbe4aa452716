#!/bin/bash
TAG=r02n
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c5.err | tee gpurun_out/${TAG}_bench_c5.json | cut -c1-200
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'k_psf|k_envelope|k_peak' -c 9 --csv --log-file gpurun_out/${TAG}_c5_counters.csv \
    python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c5.log 2>&1
