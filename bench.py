#!/usr/bin/env python
"""bench.py -- RF frames/s (and ray-segments/s) of the per-frame simulation hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): ircad11 (synthetic stand-in meshes, assets.py), santi-liver
pose, 256 scanlines x 16 Monte-Carlo samples/element, stochastic mode.  A *step* simulates
`--frames-per-step` independent frames (frame index = Philox counter) in one C-ABI call.
N > 1 (torchrun): weak scaling -- every rank simulates its own contiguous block of a freehand
probe sweep (BASELINE configs[2]) and ONE NCCL gather per step brings the RF lines to rank 0.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM and the result
left in HBM, timed with CUDA events per step (L2 flushed between steps, outside the intervals),
max over ranks.  `e2e` = the same through the host-buffer C-ABI call (poses from pinned host
memory in, RF frames into pinned host memory out, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "rf_frames_per_s"
UNIT = "frames/s"
ELEMENTS, SAMPLES = 256, 16
SCENE_REL = ("ircad11", "santi-liver.scene")


def workload_config(frames_per_step: int, n_gpus: int, gather: str = "p2p") -> dict:
    return {
        "workload": "ircad11 (synthetic organs, 624640 triangles) santi-liver pose, 256 scanlines x 16 MC samples/element, "
                    "465 RF rows, stochastic mode" + ("" if n_gpus == 1 else "; freehand probe sweep, contiguous pose blocks per rank, RF lines of every rank brought to rank 0 once per step (on its own stream: the transfer of step k overlaps the simulation of step k+1, every transfer inside a timed interval)"),
        "elements": ELEMENTS, "samples_per_element": SAMPLES, "max_depth": 10, "rf_rows": 465,
        "frames_per_step_per_gpu": frames_per_step,
        "parallelism": f"pose-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "gather": gather if n_gpus > 1 else None,
        "l2": "L2 flushed (256 MiB memset) between timed steps, outside the per-step CUDA-event intervals",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler, not OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_frames_per_s(n_frames: int, threads: int, scene_path: Path, pose: np.ndarray, first_frame: int = 0):
    from oracle import oracle_py as O
    A = O.load_scene_py(scene_path)
    osc = O.OracleScene(A)
    O.volume_raw()
    O.oracle().orc_set_threads(int(threads))
    p = O.default_params(elements=ELEMENTS, samples=SAMPLES)
    osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=first_frame)          # warm caches
    t0 = time.perf_counter()
    tests = 0
    stage = np.zeros(4)
    for f in range(n_frames):
        r = osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=first_frame + 1 + f)
        tests += r["tests"]
        stage += np.asarray(r["stage_seconds"], dtype=np.float64)[:4]
    dt = time.perf_counter() - t0
    oracle_frames_per_s.last_stage_ms_per_frame = (stage / n_frames * 1e3).tolist()      # cast, accumulate, convolve+envelope, scan
    return n_frames / dt, tests / dt, osc, p


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from mcray_tracing_b200 import assets
    from oracle import oracle_py as O
    d = assets.ensure_all()
    scene = d[SCENE_REL[0]] / SCENE_REL[1]
    pose = np.array([-17.5, 1.0, 5.0, 120.0, 0.0, -90.0], np.float32)
    threads = host_threads()
    frames_per_sample = 4
    A = O.load_scene_py(scene)
    osc = O.OracleScene(A)
    O.volume_raw()
    p = O.default_params(elements=ELEMENTS, samples=SAMPLES)
    # the reference is single-threaded (pragmas commented out, scene.cpp:74,105); the oracle's OpenMP
    # variant over elements is used when it is actually faster on this host
    best_threads, best = 1, 0.0
    for th in sorted({1, threads}):
        O.oracle().orc_set_threads(th)
        t0 = time.perf_counter()
        osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=0)
        osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=1)
        fps = 2.0 / (time.perf_counter() - t0)
        if fps > best:
            best, best_threads = fps, th
    O.oracle().orc_set_threads(best_threads)
    for w in range(args.warmup):
        osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=w)
    t0 = time.perf_counter()
    segs = 0
    for k in range(args.steps):
        for f in range(frames_per_sample):
            segs += osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=100 + k * frames_per_sample + f)["tests"]
    dt = time.perf_counter() - t0
    fps = args.steps * frames_per_sample / dt
    sample = f"{frames_per_sample} frames per step of the same workload, oracle port of the reference CPU path (reference itself needs Bullet+OpenCV, not buildable here)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.frames_per_step, world, args.gather),
        "ray_segments_per_s": segs / dt,
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": best_threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from mcray_tracing_b200 import api, assets, sweep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        d = assets.ensure_all()
    if world > 1:
        dist.barrier()
    d = assets.ensure_all()
    scene = d[SCENE_REL[0]] / SCENE_REL[1]
    F = args.frames_per_step
    params = api.default_params(elements=ELEMENTS, samples=SAMPLES)
    sim = api.Simulator(scene, params, device=local_rank)
    sim.set_option("max_batch_poses", max(F, 1))
    rows, cols = sim.rows, sim.cols
    if world == 1:
        poses = np.repeat(sim.start_pose[None, :], F, axis=0)            # configs[1]: single probe pose
        my_first = 0
    else:
        allp = assets.sweep_poses(F * world)                              # configs[2]: freehand sweep
        b, e = sweep.shard_bounds(F * world, world, rank)
        poses, my_first = allp[b:e], b
    sizes = sweep.all_shard_sizes(F * world, world)
    seed = 1234
    st = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)
    # two output buffers: the gather of step k (comm stream) overlaps the simulation of step k+1 (stream st)
    # N > 1: every rank DEPOSITS its finished RF lines straight into rank 0's double-buffered receive buffer over NVLink
    # (sweep.PeerDeposit: CUDA-IPC peer copies + a 4-byte NCCL all-reduce as the completion signal); --gather nccl selects the
    # grouped ncclSend/ncclRecv gather instead (sweep.gather_lines).  Rank 0 simulates in place in both cases.
    peer = None
    recvs = None
    if world > 1 and args.gather == "p2p":
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            peer = sweep.PeerDeposit(F, (cols, rows), dev, n_slots=2, dst=0)
        except Exception as e:                                  # e.g. no peer access: fall back on every rank
            print(f"bench.py: rank {rank}: peer deposit unavailable ({e}); falling back to the NCCL gather", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
    if peer is not None:
        if rank == 0:
            recvs = [peer.slot_tensor(b) for b in range(2)]
            outs = [r[:F] for r in recvs]
        else:
            outs = [torch.empty((F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2)]
    elif world > 1 and rank == 0:
        recvs = [torch.empty((world * F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2)]
        outs = [r[:F] for r in recvs]
    else:
        outs = [torch.empty((F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2 if world > 1 else 1)]
    out = outs[0]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    frames_per_step_total = F * world
    gather_done = [torch.cuda.Event() for _ in range(2)]
    pending = [False, False]

    def compute(k: int, i: int):
        """simulate this rank's pose block of step k into outs[i % 2] on stream st"""
        b = i % len(outs)
        if pending[b]:                                   # the gather that last read this buffer must be done
            st.wait_event(gather_done[b]); pending[b] = False
        sim.simulate_device(poses, outs[b].data_ptr(), seed=seed, first_frame=k * frames_per_step_total + my_first, stream=st.cuda_stream, sync=False)

    def gather(i: int, after: "torch.cuda.Event"):
        """bring the finished RF lines of outs[i % 2] to rank 0 on the comm stream, not before `after`"""
        b = i % len(outs)
        comm.wait_event(after)
        if peer is not None:
            peer.deposit(b, outs[b], comm)
            peer.commit(comm)
            gather_done[b].record(comm)
        else:
            with torch.cuda.stream(comm):
                sweep.gather_lines(outs[b], sizes, dst=0, out=recvs[b] if recvs is not None else None, in_place=recvs is not None)
                gather_done[b].record(comm)
        pending[b] = True

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    done_ev = torch.cuda.Event()
    for k in range(max(args.warmup, 3)):
        compute(k, k)
        if world > 1:
            done_ev.record(st)
            gather(k, done_ev)
    sync_all()
    pending[0] = pending[1] = False
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # Timed region.  Interval k = [start_k, end_k] on stream st, the L2 flush before start_k outside it.  N > 1: the
    # gather of step k-1 is released only after start_k and end_k is recorded only after st has waited for it, so
    # every gather lies entirely inside a timed interval (overlapped with the next step's simulation); the last
    # step's gather gets an interval of its own (index K).
    n_iv = args.steps + (1 if world > 1 else 0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_iv)]
    seg_total = 0
    step_total = 0
    launches = 0
    t_wall0 = time.perf_counter()
    for k in range(n_iv):
        with torch.cuda.stream(st):
            flush.fill_(float(k))                                          # evict L2 (not timed)
            ev[k][0].record(st)
        if world > 1 and k > 0:
            gather(k - 1, ev[k][0])
        if k < args.steps:
            compute(1000 + k, k)
        if world > 1 and k > 0:
            st.wait_event(gather_done[(k - 1) % 2]); pending[(k - 1) % 2] = False
        ev[k][1].record(st)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    my_ms = float(sum(ms_steps))
    stt = sim.stats()
    seg_per_step, march_per_step, launches_per_step = int(stt.segments), int(stt.march_steps), int(stt.kernel_launches)
    clock_info = clocks.stop() if rank == 0 else None
    tmax = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    segs_all = torch.tensor([seg_per_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(segs_all, op=dist.ReduceOp.SUM)
    total_ms = float(tmax.item())
    value = frames_per_step_total * args.steps / (total_ms * 1e-3)
    per_rank_ms = [my_ms / args.steps]
    if world > 1:
        allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([my_ms / args.steps], dtype=torch.float64, device=dev))
        per_rank_ms = [float(t.item()) for t in allms]

    # ---- e2e: host buffers through mcrt_simulate (pinned), copies inside the timed region ----------
    host_out = torch.empty((F, cols, rows), dtype=torch.float32, pin_memory=True)
    host_np = host_out.numpy()
    for k in range(3):
        sim.simulate(poses, seed=seed, first_frame=k * frames_per_step_total + my_first, rf_out=host_np)
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sim.simulate(poses, seed=seed, first_frame=(2000 + k) * frames_per_step_total + my_first, rf_out=host_np)
        checksum = float(host_np[0, cols // 2, rows // 2])                 # the step's result is read on the host
    e2e_sync_s = time.perf_counter() - t0
    # the same through the streaming driver (stream.FrameStreamer, public API): every step still uploads its
    # poses from host memory and lands its RF frames in pinned host memory, but the PCIe copy of step k
    # overlaps the simulation of step k+1 (separate copy stream, ring of 3 pinned buffers)
    from mcray_tracing_b200 import stream as mstream
    fs = mstream.FrameStreamer(sim, depth=3, frames_per_submit=F, seed=seed)
    fs._frame = 4000 * frames_per_step_total + my_first
    for k in range(3):
        fs.submit(poses)
        fs._frame += frames_per_step_total - F
    while fs.pending():
        fs.get()
    sync_all()
    t0 = time.perf_counter()
    submitted = 0
    checksum = 0.0
    while submitted < args.steps or fs.pending():
        while submitted < args.steps and fs._free:
            fs.submit(poses)
            fs._frame += frames_per_step_total - F        # global frame index advances by the whole job's step
            submitted += 1
        _, rf_host, _ = fs.get()
        checksum += float(rf_host[0, cols // 2, rows // 2])               # the step's result is read on the host
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step_total * args.steps / float(e2e_t[0].item())
    e2e_sync_value = frames_per_step_total * args.steps / float(e2e_t[1].item())

    # ---- latency mode (one frame per call) and per-stage kernel times (rank 0 only) ---------------
    extra = {}
    roofline = None
    cpu_baseline = None
    if rank == 0:
        one = poses[:1]
        out1 = out[:1]
        for k in range(5):
            sim.simulate_device(one, out1.data_ptr(), seed=seed, first_frame=k)
        lat = []
        for k in range(20):
            flush.fill_(0.0)
            torch.cuda.synchronize(dev)
            sim.simulate_device(one, out1.data_ptr(), seed=seed, first_frame=50 + k)
            lat.append(sim.stats().ms_total)
        extra["latency_mode"] = {"frames_per_call": 1, "ms_per_frame_device": float(np.median(lat)), "frames_per_s": 1e3 / float(np.median(lat)),
                                 "l2": "flushed before every frame (BVH, volume and segments come from HBM)"}
        warm = []
        for k in range(20):
            sim.simulate_device(one, out1.data_ptr(), seed=seed, first_frame=100 + k)
            warm.append(sim.stats().ms_total)
        extra["latency_mode"]["ms_per_frame_device_warm_l2"] = float(np.median(warm))
        extra["latency_mode"]["frames_per_s_warm_l2"] = 1e3 / float(np.median(warm))
        if world == 1:
            # the N > 1 runs simulate a freehand sweep (BASELINE configs[2]); its per-frame cost differs from
            # the single pose, so the 1-GPU figure on the SAME sweep poses is reported for a like-for-like
            # scaling comparison
            sw = assets.sweep_poses(F * 8)[:F]
            ms = []
            for k in range(3 + 8):
                flush.fill_(0.0)
                torch.cuda.synchronize(dev)
                sim.simulate_device(sw, out.data_ptr(), seed=seed, first_frame=k * F)
                if k >= 3:
                    ms.append(sim.stats().ms_total)
            extra["sweep_workload_1gpu"] = {"frames_per_step": F, "poses": "first rank's block of an 8-rank sweep", "ms_per_step": float(np.mean(ms)),
                                            "frames_per_s": F / (float(np.mean(ms)) * 1e-3)}
        # per-stage device times: separate pass, stage events between the kernels (no CUDA graph)
        sim.set_option("profile_stages", 1)
        tr, ac, po, tot, msteps = [], [], [], [], []
        for k in range(2 + 5):
            flush.fill_(0.0)
            torch.cuda.synchronize(dev)
            sim.simulate_device(poses, out.data_ptr(), seed=seed, first_frame=(3000 + k) * frames_per_step_total + my_first)
            s = sim.stats()
            if k >= 2:
                tr.append(s.ms_trace); ac.append(s.ms_accumulate); po.append(s.ms_post); tot.append(s.ms_total); msteps.append(s.march_steps)
        sim.set_option("profile_stages", 0)
        ms_acc = float(np.mean(ac))
        alg_bytes = 8.0 * float(np.mean(msteps)) + 4.0 * F * cols * rows      # 8 B / march step + 4 B / RF sample
        peak, peak_src = measured_hbm_peak()
        achieved = alg_bytes / (ms_acc * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "accumulate_traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                # the ncu capture was taken at frames_per_launch frames per call; DRAM bytes scale with the frames of a launch
                traffic = int(tj["dram_bytes_per_launch"] * F / tj.get("frames_per_launch", F))
            except Exception:
                traffic = None
        roofline = {"kernel": "k_accumulate_win (echo accumulation + scatterer-volume gather + sample reduction, one kernel)", "bound": "hbm", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_acc,
                    "stage_ms": {"trace": float(np.mean(tr)), "accumulate": ms_acc, "post": float(np.mean(po)), "total": float(np.mean(tot))}}
        # the trace stage (10 k_bounce launches) is the larger share of the step but is SM-issue / latency bound, not
        # bandwidth bound: SURVEY.md 8(d) counts 32 B ray in + 16 B hit out per closest-hit query
        ms_tr = float(np.mean(tr))
        tr_bytes = 48.0 * seg_per_step
        extra["roofline_trace"] = {"kernel": "k_bounce x max_depth (BVH traversal + boundary physics + compaction)", "bound": "hbm",
                                   "achieved": tr_bytes / (ms_tr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": tr_bytes / (ms_tr * 1e-3) / 1e9 / peak,
                                   "algorithmic_bytes_per_step": tr_bytes, "stage_ms": ms_tr, "closest_hit_queries_per_s": seg_per_step / (ms_tr * 1e-3),
                                   "note": "not bandwidth-bound by construction (48 B per query): limited by issue slots and dependent node fetches, "
                                           "see profiles/ (smsp__issue_active ~57 %, 19-26 of 32 lanes active per instruction)"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle_py as O
            threads = host_threads()
            n = args.cpu_frames
            fps1, sps1, _, _ = oracle_frames_per_s(n, 1, scene, sim.start_pose)
            stage1 = list(getattr(oracle_frames_per_s, "last_stage_ms_per_frame", []))
            fpsN, spsN = (fps1, sps1) if threads == 1 else oracle_frames_per_s(n, threads, scene, sim.start_pose)[:2]
            stageN = list(getattr(oracle_frames_per_s, "last_stage_ms_per_frame", []))
            best_fps, best_sps, cores = (fpsN, spsN, threads) if fpsN > fps1 else (fps1, sps1, 1)
            cpu_baseline = {"value": best_fps, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} frames of the same workload per variant; single thread (as the reference ships, scene.cpp:74) = {fps1:.2f} frames/s, "
                                      f"OpenMP over elements on {threads} threads = {fpsN:.2f} frames/s; oracle port (the reference needs Bullet+OpenCV)",
                            "ray_segments_per_s": best_sps,
                            "single_thread": {"frames_per_s": fps1, "ray_segments_per_s": sps1, "stage_ms_per_frame": stage1},
                            "all_threads": {"threads": threads, "frames_per_s": fpsN, "ray_segments_per_s": spsN, "stage_ms_per_frame": stageN},
                            "stage_order": ["cast_rays", "accumulate", "convolve+envelope", "scan_convert"]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(F, world, ("p2p deposit over NVLink (CUDA IPC) + 4-byte NCCL all-reduce" if peer is not None else "NCCL grouped send/recv") if world > 1 else "p2p"),
            "ray_segments_per_s": float(segs_all.item()) * args.steps / (total_ms * 1e-3),
            "segments_per_step": float(segs_all.item()), "march_steps_per_step_per_gpu": march_per_step,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * 24 + 16), "d2h_bytes_per_step": int(F * cols * rows * 4),
                    "timing": "wall clock around K steps through stream.FrameStreamer (mcrt_simulate_async + pinned ring of 3 buffers): "
                              "per step poses host->device and RF frames device->pinned host, the copy of step k overlapping step k+1",
                    "synchronous_call_value": e2e_sync_value,
                    "synchronous_call_timing": "wall clock around K blocking mcrt_simulate calls with pinned host buffers (no overlap)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "wall_s_timed_region": t_wall,
            "ms_per_step_per_rank": per_rank_ms,      # the sweep's pose blocks differ in cost; value uses the slowest rank
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if peer is not None:
        peer.close()
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=512,
                    help="independent frames per C-ABI call and per GPU (measured: 256 -> 81.7k, 512 -> 87.0k, 1024 -> 89.4k frames/s)")
    ap.add_argument("--cpu-frames", type=int, default=60, help="frames per CPU-baseline variant (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N > 1: peer-memory deposit over NVLink (default) or NCCL send/recv gather")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
