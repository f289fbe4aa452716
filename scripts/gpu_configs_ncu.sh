#!/bin/bash
# ncu evidence at BASELINE configs 4 and 5 (one GPU).  usage: scripts/gpu_configs_ncu.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python scripts/bench_configs.py --out gpurun_out/${TAG}_configs.json 2>&1 | tail -3 | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:"k_bounce" -s 20 -c 10 -f -o gpurun_out/${TAG}c4_prof \
    python scripts/bench_configs.py --only c4 --out gpurun_out/tmp_c4.json > gpurun_out/${TAG}_c4_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_reduce|k_psf|k_envelope|k_post|k_peak|k_env" -s 8 -c 8 -f -o gpurun_out/${TAG}c5_prof \
    python scripts/bench_configs.py --only c5 --out gpurun_out/tmp_c5.json > gpurun_out/${TAG}_c5_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG}
