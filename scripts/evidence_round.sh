#!/bin/bash
# One gpurun call worth of round evidence: GPU test suite, bench lines of every configuration + the reference arm, ncu launch
# list, per-stage instruction counters, full-set captures of the three stage kernels.
# usage: scripts/evidence_round.sh <tag> [skip-tests]
TAG=${1:-r02}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
fi
python bench.py --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-200
for c in c3 c4 c5; do
  python bench.py --config $c --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_$c.err | tee gpurun_out/${TAG}_bench_$c.json | cut -c1-300
done
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# per-stage instruction counters (one step of each configuration that reports issue fractions)
for c in c2 c4 c5; do
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -c 150 --csv \
      --log-file gpurun_out/${TAG}_counters_$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_counters_$c.log 2>&1
done
# the three stage kernels, full set (one launch each of the steady state; k_bounce: bounces 1-3)
ncu --set full --clock-control none --import-source on -k regex:"k_bounce|k_accumulate_win|k_post_tma|k_first_hit" -s 24 -c 7 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
