#!/bin/bash
# r02an: final state check -- smoke(), the whole GPU suite, one bench line
TAG=r02an
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-400
