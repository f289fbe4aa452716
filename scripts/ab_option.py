"""A/B of a runtime option (development aid): stage times at 256 frames per call, ircad11 256x16.
usage: ab_option.py option=v0,v1 [frames]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from mcray_tracing_b200 import api, assets

opt, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = assets.ensure_all()
sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
sim.set_option("max_batch_poses", F)
poses = np.repeat(sim.start_pose[None, :], F, axis=0)
out = torch.empty((F, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for rep in range(2):
    for v in vals:
        sim.set_option(opt, v)
        sim.set_option("profile_stages", 1)
        acc = []
        for k in range(8):
            flush.fill_(0.0); torch.cuda.synchronize()
            sim.simulate_device(poses, out.data_ptr(), seed=1234, first_frame=k * F)
            s = sim.stats()
            if k >= 2:
                acc.append((s.ms_trace, s.ms_accumulate, s.ms_post, s.ms_total))
        sim.set_option("profile_stages", 0)
        tot = []
        for k in range(8):
            flush.fill_(0.0); torch.cuda.synchronize()
            sim.simulate_device(poses, out.data_ptr(), seed=1234, first_frame=k * F)
            if k >= 2:
                tot.append(sim.stats().ms_total)
        a = np.mean(acc, axis=0)
        print(f"{opt}={v}: trace {a[0]:.3f} accumulate {a[1]:.3f} post {a[2]:.3f} total(staged) {a[3]:.3f}  graph total {np.mean(tot):.3f} ms -> {F/np.mean(tot)*1e3:.0f} frames/s")
