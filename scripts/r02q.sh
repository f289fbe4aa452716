#!/bin/bash
TAG=r02q
mkdir -p gpurun_out
python scripts/ab_option.py overlap=0,1 512 2>&1 | tee gpurun_out/${TAG}_ab_overlap.txt
python scripts/ab_option.py overlap=0,1 1024 2>&1 | tee -a gpurun_out/${TAG}_ab_overlap.txt
python scripts/ab_option.py bvh_builder=0,1 512 2>&1 | tee gpurun_out/${TAG}_ab_sah.txt
