"""A/B of library builds (development aid): for the library named by MCRT_LIB_PATH prints, on ircad11 256x16 at F frames
per call, the stage times, frames/s and the traversal work counters; with --c4 the same on the 2M-triangle stress scene.
usage: MCRT_LIB_PATH=... python scripts/ab_libs.py [frames] [--c4] [--opt name=value ...]"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from mcray_tracing_b200 import api, assets

args = [a for a in sys.argv[1:] if not a.startswith("--")]
F = int(args[0]) if args else 512
opts = []
for i, a in enumerate(sys.argv):
    if a == "--opt":
        k, v = sys.argv[i + 1].split("=")
        opts.append((k, int(v)))
tag = os.path.basename(os.environ.get("MCRT_LIB_PATH", "libmcrt.so")) + "".join(f" {k}={v}" for k, v in opts)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def run(sim, poses, label):
    n = len(poses)
    out = torch.empty((n, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
    sim.set_option("max_batch_poses", n)
    for k, v in opts:
        sim.set_option(k, v)
    try:
        sim.set_option("bvh_wait", 1)          # (older A/B libraries do not know the option)
    except Exception:
        pass
    sim.set_option("profile_stages", 1)
    acc = []
    for k in range(7):
        flush.fill_(0.0); torch.cuda.synchronize()
        sim.simulate_device(poses, out.data_ptr(), seed=1234, first_frame=k * n)
        s = sim.stats()
        if k >= 2:
            acc.append((s.ms_trace, s.ms_accumulate, s.ms_post, s.ms_total))
    sim.set_option("profile_stages", 0)
    tot = []
    for k in range(7):
        flush.fill_(0.0); torch.cuda.synchronize()
        sim.simulate_device(poses, out.data_ptr(), seed=1234, first_frame=k * n)
        if k >= 2:
            tot.append(sim.stats().ms_total)
    sim.set_option("count_traversal", 1)
    sim.simulate_device(poses, out.data_ptr(), seed=1234, first_frame=0)
    st = sim.stats()
    sim.set_option("count_traversal", 0)
    a = np.mean(acc, axis=0)
    print(f"[{tag}] {label} F={n}: trace {a[0]:.3f} accumulate {a[1]:.3f} post {a[2]:.3f} | graph total {np.mean(tot):.3f} ms -> {n / np.mean(tot) * 1e3:.0f} frames/s"
          f" | segments {st.segments} node visits/seg {st.bvh_node_visits / st.segments:.2f} tri tests/seg {st.bvh_triangle_tests / st.segments:.2f}", flush=True)


d = assets.ensure_all()
sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
run(sim, np.repeat(sim.start_pose[None, :], F, axis=0), "ircad11")
run(sim, np.repeat(sim.start_pose[None, :], 1, axis=0), "ircad11")
sim.close()
if "--c4" in sys.argv:
    A = assets.stress_scene_arrays()
    sim = api.Simulator(A, api.default_params(elements=512, samples=16))
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])[None, :]
    run(sim, np.repeat(pose, 64, axis=0), "C4 stress 2M tris")
    sim.close()
