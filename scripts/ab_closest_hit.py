"""Traversal micro-benchmark (development aid): closest-hit kernel alone on the rays of a real wavefront
(ircad11 256x16, 64 poses), grouped by bounce like the wavefront issues them.  usage: ab_closest_hit.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from mcray_tracing_b200 import api, assets

d = assets.ensure_all()
sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
fr, to = [], []
per_depth = [[] for _ in range(10)]
for f in range(48):
    segs, n = sim.cast_rays(sim.start_pose, seed=1234, frame=f)
    D = segs.shape[-1]
    for k in range(D):
        v = n > k
        s = segs[..., k][v]
        per_depth[k].append((s["from"] + 0.1 * s["dir"], s["from"] + 40.0 * s["dir"]))
F = np.concatenate([np.concatenate([a for a, _ in per_depth[k]]) for k in range(10) if per_depth[k]]).astype(np.float32)
T = np.concatenate([np.concatenate([b for _, b in per_depth[k]]) for k in range(10) if per_depth[k]]).astype(np.float32)
ms = []
for it in range(6):
    sim.closest_hit(F, T)
    ms.append(sim.stats().ms_total)
print(f"rays {len(F)}  closest-hit kernel ms {np.median(ms[1:]):.3f}  -> {len(F)/np.median(ms[1:])*1e-6:.2f} G rays/s")
