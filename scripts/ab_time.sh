#!/bin/bash
# A/B timing of library variants (development aid)
for lib in "$@"; do
  echo "== $lib"
  MCRT_LIB_PATH=$PWD/mcray_tracing_b200/$lib python scripts/quick_time.py 2>&1 | grep -E "stages|batch 64"
done
