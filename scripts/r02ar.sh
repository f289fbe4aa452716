#!/bin/bash
TAG=r02ar
mkdir -p gpurun_out
(python scripts/ab_option.py tail_merge=1,2,3,4 1024; python scripts/ab_option.py tail_merge=1,2,3,4 256; python scripts/ab_option.py tail_merge=1,2,3,4 64;  python scripts/ab_option.py tail_merge=1,2,3,4 16) 2>&1 | tee gpurun_out/${TAG}_ab_tail.txt
