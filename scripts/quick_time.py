"""Quick timing probe (development aid): ircad11 256x16, latency mode and batched."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from mcray_tracing_b200 import api, assets

d = assets.ensure_all()
scene = d["ircad11"] / "santi-liver.scene"
t0 = time.time()
sim = api.Simulator(scene, api.default_params(elements=256, samples=16))
print("create s", time.time() - t0, "tris", sim.info.n_triangles, "nodes", sim.info.n_bvh_nodes)
pose = sim.start_pose[None, :]
for nb in (1, 8, 64, 256):
    poses = np.repeat(pose, nb, axis=0)
    for it in range(3):
        rf = sim.simulate(poses, seed=1, first_frame=it * nb)
    st = sim.stats()
    print(f"batch {nb}: ms_total {st.ms_total:.3f}  per-frame {st.ms_total/nb*1000:.1f} us  fps {nb/st.ms_total*1000:.0f}  segs {st.segments} steps {st.march_steps} launches {st.kernel_launches}")
sim.set_option("profile_stages", 1)
for nb in (1, 64):
    poses = np.repeat(pose, nb, axis=0)
    for it in range(2):
        sim.simulate(poses, seed=1, first_frame=0)
    st = sim.stats()
    print(f"stages batch {nb}: trace {st.ms_trace:.3f} acc {st.ms_accumulate:.3f} post {st.ms_post:.3f} total {st.ms_total:.3f}")
sim.set_option("profile_stages", 0)
sim.set_option("count_traversal", 1)
poses = np.repeat(pose, 64, axis=0)
sim.simulate(poses, seed=1, first_frame=0)
st = sim.stats()
print(f"traversal: segs {st.segments} node visits/seg {st.bvh_node_visits/st.segments:.1f} tri tests/seg {st.bvh_triangle_tests/st.segments:.1f}")
# SAH tree experiment
import time as _t
t0 = _t.time(); sim.set_option("bvh_builder", 1); print("sah build s", _t.time() - t0)
sim.simulate(poses, seed=1, first_frame=0)
st = sim.stats()
print(f"SAH traversal: segs {st.segments} node visits/seg {st.bvh_node_visits/st.segments:.1f} tri tests/seg {st.bvh_triangle_tests/st.segments:.1f}")
sim.set_option("count_traversal", 0)
sim.set_option("profile_stages", 1)
for it in range(2):
    sim.simulate(poses, seed=1, first_frame=0)
st = sim.stats()
print(f"SAH stages batch 64: trace {st.ms_trace:.3f} acc {st.ms_accumulate:.3f} post {st.ms_post:.3f} total {st.ms_total:.3f}")
