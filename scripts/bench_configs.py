#!/usr/bin/env python
"""Measurements at the other BASELINE configurations (C3 sweep, C4 traversal stress, C5 bandwidth
stress) on one GPU.  Not the headline bench (bench.py); results are committed under profiles/.

    python scripts/bench_configs.py [--only c3,c4,c5] [--out gpurun_out/configs.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mcray_tracing_b200 import api, assets  # noqa: E402


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def timed(sim, poses, out, reps=5, warm=2, seed=1):
    ms = []
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for k in range(warm + reps):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        sim.simulate_device(poses, out.data_ptr(), seed=seed, first_frame=k * len(poses))
        if k >= warm:
            ms.append(sim.stats().ms_total)
    return float(np.mean(ms)), sim.stats()


def stages(sim, poses, out, reps=3, seed=1):
    sim.set_option("profile_stages", 1)
    acc = []
    for k in range(1 + reps):
        torch.cuda.synchronize()
        sim.simulate_device(poses, out.data_ptr(), seed=seed, first_frame=k * len(poses))
        s = sim.stats()
        if k >= 1:
            acc.append((s.ms_trace, s.ms_accumulate, s.ms_post, s.ms_total))
    sim.set_option("profile_stages", 0)
    a = np.mean(np.array(acc), axis=0)
    return dict(trace=float(a[0]), accumulate=float(a[1]), post=float(a[2]), total=float(a[3]))


def c3():
    d = assets.ensure_all()
    sim = api.Simulator(d["ircad11"] / "santi-liver.scene", api.default_params(elements=256, samples=16))
    poses = assets.sweep_poses(512)
    sim.set_option("max_batch_poses", 128)
    out = torch.empty((512, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
    ms, st = timed(sim, poses, out, reps=3, warm=1)
    r = dict(config="C3: ircad11 512-pose freehand sweep, 256x16, one GPU, batches of 128 poses", ms_per_sweep=ms, frames_per_s=512 / ms * 1e3,
             segments_per_s=st.segments / ms * 1e3, segments=int(st.segments), march_steps=int(st.march_steps))
    sim.close()
    return r


def c4():
    A = assets.stress_scene_arrays()
    sim = api.Simulator(A, api.default_params(elements=512, samples=16))
    pose = np.concatenate([A["transducer_position"], A["transducer_angles"]])[None, :]
    poses = np.repeat(pose, 8, axis=0)
    out = torch.empty((8, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
    ms, st = timed(sim, poses, out)
    stg = stages(sim, poses, out)
    sim.set_option("count_traversal", 1)
    sim.simulate_device(poses, out.data_ptr(), seed=1, first_frame=0)
    sc = sim.stats()
    r = dict(config="C4: 2 097 152-triangle nested shells, shininess 2 / thickness 0.5, 512 el x 16 samples x 10 bounces, 8 frames per call",
             triangles=int(sim.info.n_triangles), ms_per_call=ms, frames_per_s=8 / ms * 1e3, segments=int(st.segments),
             segments_per_s=st.segments / ms * 1e3, trace_segments_per_s=st.segments / stg["trace"] * 1e3, stage_ms=stg,
             mean_segments_per_path=st.segments / (8 * 512 * 16), bvh_node_visits_per_segment=sc.bvh_node_visits / sc.segments,
             triangle_tests_per_segment=sc.bvh_triangle_tests / sc.segments)
    # the same with the wavefront radix-sorted by origin Morton code + direction octant between bounces
    sim.set_option("count_traversal", 0)
    sim.set_option("coherence_sort", 1)
    ms2, st2 = timed(sim, poses, out)
    stg2 = stages(sim, poses, out)
    r["coherence_sort"] = dict(ms_per_call=ms2, frames_per_s=8 / ms2 * 1e3, segments_per_s=st2.segments / ms2 * 1e3,
                               trace_segments_per_s=st2.segments / stg2["trace"] * 1e3, stage_ms=stg2)
    # a call large enough to fill the machine (64 frames = 524 288 paths): plain, and with the coherence sort
    sim.set_option("coherence_sort", 0)
    poses64 = np.repeat(pose, 64, axis=0)
    out64 = torch.empty((64, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
    sim.set_option("max_batch_poses", 64)
    for name, opt in (("plain", 0), ("coherence_sort", 1)):
        sim.set_option("coherence_sort", opt)
        ms3, st3 = timed(sim, poses64, out64, reps=3, warm=1)
        stg3 = stages(sim, poses64, out64)
        r["frames_per_call_64_" + name] = dict(ms_per_call=ms3, frames_per_s=64 / ms3 * 1e3, segments=int(st3.segments),
                                               segments_per_s=st3.segments / ms3 * 1e3, trace_segments_per_s=st3.segments / stg3["trace"] * 1e3,
                                               stage_ms=stg3)
    sim.close()
    return r


def c5():
    d = assets.ensure_all()
    peak = hbm_peak()
    res = []
    for ka, kl in ((63, 31), (31, 15)):
        p = api.default_params(elements=1024, samples=16, axial_scale=17.6, psf_axial=ka, psf_lateral=kl)
        sim = api.Simulator(d["ircad11"] / "santi-liver.scene", p)
        pose = sim.start_pose[None, :]
        NF = 8
        poses = np.repeat(pose, NF, axis=0)
        out = torch.empty((NF, sim.cols, sim.rows), dtype=torch.float32, device="cuda")
        ms, st = timed(sim, poses, out, reps=3, warm=1)
        stg = stages(sim, poses, out)
        px = NF * sim.cols * sim.rows
        steps = st.march_steps
        acc_alg = 8.0 * steps + 4.0 * px
        res.append(dict(psf=f"{ka}x{kl}", rows=sim.rows, cols=sim.cols, frames_per_call=NF, ms_per_call=ms, stage_ms=stg, march_steps=int(steps),
                        accumulate_algorithmic_GBps=acc_alg / stg["accumulate"] / 1e6, accumulate_frac_of_measured_hbm=acc_alg / stg["accumulate"] / 1e6 / peak,
                        post_pixels=px, post_algorithmic_bytes=3 * 8.0 * px, post_GBps=3 * 8.0 * px / stg["post"] / 1e6,
                        post_frac_of_measured_hbm=3 * 8.0 * px / stg["post"] / 1e6 / peak,
                        post_flops_per_pixel=2 * (ka + kl), post_GFLOPs=2.0 * (ka + kl) * px / stg["post"] / 1e6))
        sim.close()
    return dict(config="C5: 1024 scanlines x 8333 samples (axial_scale 17.6), 16 samples/element, large PSF; post = axial + lateral + envelope "
                       "(3 passes x 8 B/pixel)", hbm_peak_GBps=peak, runs=res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c3,c4,c5")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "configs.json"))
    a = ap.parse_args()
    out = {}
    for name, fn in (("c3", c3), ("c4", c4), ("c5", c5)):
        if name in a.only.split(","):
            t0 = time.time()
            out[name] = fn()
            out[name]["wall_s"] = time.time() - t0
            print(name, json.dumps(out[name]), flush=True)
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
