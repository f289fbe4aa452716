#!/bin/bash
# r02am: how good is a SAH tree whose splits are restricted to prefixes of the Morton order (what a device builder could do without moving data)?
TAG=r02am
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulate_matches or windowed" 2>&1 | tail -2
(echo "# bvh_builder=0 device LBVH (Karras), 1 host binned SAH; second block: MCRT_SAH_MODE=morton makes builder 1 a sweep SAH over the Morton order";
 python scripts/ab_option.py bvh_builder=0,1 512; echo "# MCRT_SAH_MODE=morton"; MCRT_SAH_MODE=morton python scripts/ab_option.py bvh_builder=0,1 512) 2>&1 | tee gpurun_out/${TAG}_ab_morton_sweep_sah.txt
python - <<'PY' 2>&1 | tee -a gpurun_out/r02am_ab_morton_sweep_sah.txt
import os, sys, time; sys.path.insert(0, '.')
import numpy as np
from mcray_tracing_b200 import api, assets
d = assets.ensure_all()
for mode in ("", "morton"):
    if mode: os.environ["MCRT_SAH_MODE"] = mode
    sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
    poses = np.repeat(sim.start_pose[None, :], 64, axis=0)
    for b in (0, 1):
        t0 = time.time(); sim.set_option("bvh_builder", b); dt = time.time() - t0
        sim.set_option("count_traversal", 1)
        sim.simulate(poses, seed=1, first_frame=0); st = sim.stats()
        sim.set_option("count_traversal", 0)
        print(f"mode '{mode}' builder {b}: build {dt*1e3:.1f} ms, node visits/seg {st.bvh_node_visits/st.segments:.2f} tri tests/seg {st.bvh_triangle_tests/st.segments:.2f}")
    sim.close()
PY
