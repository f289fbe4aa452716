#!/bin/bash
TAG=r02r
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "traversal_options or moving" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
python scripts/ab_option.py bvh_builder=0,1,2 512 2>&1 | tee gpurun_out/${TAG}_ab_ploc.txt
python - <<'PY' 2>&1 | tee -a gpurun_out/${TAG}_ab_ploc.txt
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from mcray_tracing_b200 import api, assets
d = assets.ensure_all()
sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
poses = np.repeat(sim.start_pose[None, :], 64, axis=0)
for b in (0, 1, 2):
    t0 = time.time(); sim.set_option("bvh_builder", b); dt = time.time() - t0
    sim.set_option("count_traversal", 1)
    sim.simulate(poses, seed=1, first_frame=0); st = sim.stats()
    sim.set_option("count_traversal", 0)
    print(f"builder {b}: build {dt*1e3:.1f} ms, node visits/seg {st.bvh_node_visits/st.segments:.2f} tri tests/seg {st.bvh_triangle_tests/st.segments:.2f}")
PY
