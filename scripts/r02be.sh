#!/bin/bash
# r02be: last commit of the round -- smoke, the optimiser / streaming / mesh-update tests, one bench line
TAG=r02be
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
python -m pytest tests -m gpu -x -q -k "background or streaming or moving or traversal_options or different_streams" 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-200
