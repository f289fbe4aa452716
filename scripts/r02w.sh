#!/bin/bash
# r02w: long-scanline post kernels with compile-time taps + shared-memory envelope (config 5)
TAG=r02w
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "convolve_and_envelope or envelope_long or depth_dependent" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "depth_dependent or scanline_block or c5 or long" 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_pytest.log
for v in 1 0 1 0; do
  python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline --option long_ct=$v 2> gpurun_out/${TAG}_c5_$v.err | tee gpurun_out/${TAG}_bench_c5_ct$v.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
r = b['roofline']
print('long_ct=$v', 'ms/step', round(b['ms_per_step'], 3), 'stages', {k: round(v, 3) for k, v in r['stage_ms'].items()}, 'frac', round(r['frac_of_max_bytes_flops_roof'], 3))
"
done
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:"k_psf|k_envelope" -c 9 --csv \
    --log-file gpurun_out/${TAG}_c5_counters.csv python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c5.log 2>&1
python - <<'PY'
import csv
lines = [l for l in open('gpurun_out/r02w_c5_counters.csv') if l.startswith('"')]
d = {}
for r in csv.DictReader(lines):
    d.setdefault(int(r['ID']), {'k': r['Kernel Name'][:40]})[r['Metric Name']] = r['Metric Value']
for i in sorted(d)[:6]:
    print(d[i])
PY
