#!/bin/bash
TAG=r02as
mkdir -p gpurun_out
(python scripts/ab_option.py tail_merge=3,5,8,12,20,40 128; python scripts/ab_option.py tail_merge=3,5,8,12,20,40 256; python scripts/ab_option.py tail_merge=3,8,20,40 1024; python scripts/ab_option.py tail_merge=1,3,5,8 32) 2>&1 | tee gpurun_out/${TAG}_ab_tail.txt
