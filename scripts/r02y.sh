#!/bin/bash
# r02y: ray-tree mode without a sort and without HBM columns
TAG=r02y
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "ray_tree" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "ray_tree and rough" 2>&1 | tail -8 | tee gpurun_out/${TAG}_memcheck.log
python scripts/ab_option.py ray_tree=0,64 32 2>&1 | tee gpurun_out/${TAG}_ab_raytree.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_pytest.log
