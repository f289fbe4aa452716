#!/bin/bash
# One gpurun call worth of measurement: parity tests, bench line, ncu launch list, ncu full capture.
# usage: scripts/gpu_round.sh <tag> [kernel-regex]
TAG=${1:-r01}
KREGEX=${2:-k_accumulate}
KCOUNT=${3:-2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference.json
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# the top kernel, full set
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 2 -c ${KCOUNT} -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
