#!/usr/bin/env python
"""bench.py -- RF frames/s (and ray-segments/s) of the per-frame simulation hot path.

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)
    python bench.py --config c3|c4|c5 ...                            # the other BASELINE.json configurations

Workloads (BASELINE.json `configs`):
  c2 (default, the configuration the metric is quoted on): ircad11 (synthetic stand-in meshes, assets.py), santi-liver pose,
      256 scanlines x 16 Monte-Carlo samples/element, stochastic mode.  N > 1 (torchrun): weak scaling -- every rank simulates
      its share of a freehand probe sweep, poses dealt out round-robin, finished RF lines deposited on rank 0 once per step.
  c3: the 512-pose freehand sweep as ONE job (strong scaling: 512 / N poses per rank and step).
  c4: synthetic 2 097 152-triangle nested-shell mesh, 512 x 16, rough surfaces (traversal / divergence stress).
  c5: 1024 scanlines x 8333 RF rows (the reference's row formula, rfimage.h:180, cannot give exactly 8192), 63 x 31 PSF,
      plus the post-processing kernels alone on a synthetic 1024 x 8192 image.
A *step* simulates `frames_per_step` independent frames (frame index = Philox counter) in one C-ABI call.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM and the result left in HBM, timed with CUDA
events per step (L2 flushed between steps, outside the intervals), max over ranks.  `e2e` = the same through the host-buffer
C-ABI path (poses from host memory in, RF frames into pinned host memory out, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import hashlib
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


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def workload(name: str, frames_per_step: int | None):
    """-> dict(scene=path | arrays, params=dict, F=frames per step per GPU, label, sweep=bool (poses from the freehand sweep))"""
    from mcray_tracing_b200 import assets
    d = assets.ensure_all()
    if name in ("c2", "c3"):
        w = dict(scene=d["ircad11"] / "santi-liver.scene", params=dict(elements=256, samples=16), F=frames_per_step or (512 if name == "c3" else 1024),
                 label="ircad11 (synthetic organs, 624640 triangles), 256 scanlines x 16 MC samples/element, 465 RF rows, stochastic mode",
                 sweep=(name == "c3"), pose=None)
        w["label"] += ("; 512-pose freehand probe sweep (BASELINE configs[2])" if name == "c3" else "; santi-liver pose (BASELINE configs[1])")
        return w
    if name == "c4":
        A = assets.stress_scene_arrays()
        pose = np.concatenate([A["transducer_position"], A["transducer_angles"]]).astype(np.float32)
        return dict(scene=A, params=dict(elements=512, samples=16), F=frames_per_step or 128, sweep=False, pose=pose,
                    label="synthetic 2 097 152-triangle nested-shell tissue mesh, shininess 2 / thickness 0.5 on every material (rough), "
                          "512 scanlines x 16 MC samples x 10 bounces (BASELINE configs[3])")
    if name == "c5":
        return dict(scene=d["ircad11"] / "santi-liver.scene", F=frames_per_step or 32, sweep=False, pose=None,
                    params=dict(elements=1024, samples=16, axial_scale=17.6, psf_axial=63, psf_lateral=31),
                    label="ircad11 at 1024 scanlines x 8333 RF rows (axial_scale 17.6: 18 um rows; the reference's integer row formula "
                          "rfimage.h:180 has no setting that gives exactly 8192), 16 MC samples/element, 63 x 31-tap PSF (BASELINE configs[4])")
    raise SystemExit(f"bench.py: unknown --config {name}")


def workload_config(name: str, w: dict, n_gpus: int, parallelism: str, gather: str | None) -> dict:
    p = w["params"]
    return {"workload": w["label"], "config": name, "elements": p["elements"], "samples_per_element": p["samples"], "max_depth": 10,
            "frames_per_step_per_gpu": w["F"], "parallelism": parallelism, "gather": gather,
            "l2": "L2 flushed (256 MiB memset) between timed steps, outside the per-step CUDA-event intervals"}


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


def bind_to_gpu_numa_node(gpu_index: int) -> str:
    """Pin this process to the CPUs next to its GPU BEFORE it allocates pinned host buffers (first touch puts the pages on that
    NUMA node): with 8 ranks landing 244 MB per step each, buffers on the wrong socket halve the host-side rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {gpu_index}"
        return "no narrower affinity available"
    except Exception as e:                        # best effort
        return f"unavailable ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_scene(w: dict):
    from oracle import oracle_py as O
    A = w["scene"] if isinstance(w["scene"], dict) else O.load_scene_py(w["scene"])
    osc = O.OracleScene(A)
    O.volume_raw()
    return O, osc, O.default_params(**w["params"])


def oracle_frames_per_s(w: dict, pose: np.ndarray, n_frames: int, threads: int, budget_s: float = 25.0):
    """frames/s and segments/s of the oracle on `threads` host threads over at most n_frames frames / budget_s seconds"""
    O, osc, p = oracle_scene(w)
    O.oracle().orc_set_threads(int(threads))
    osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=0)          # warm caches
    t0 = time.perf_counter()
    tests, done = 0, 0
    stage = np.zeros(4)
    while done < n_frames and (done == 0 or time.perf_counter() - t0 < budget_s):
        r = osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=1 + done)
        tests += r["tests"]
        stage += np.asarray(r["stage_seconds"], dtype=np.float64)[:4]
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, tests / dt, (stage / done * 1e3).tolist(), done


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    w = workload(args.config, args.frames_per_step)
    from mcray_tracing_b200 import assets
    pose = w["pose"] if w["pose"] is not None else np.array([-17.5, 1.0, 5.0, 120.0, 0.0, -90.0], np.float32)
    threads = host_threads()
    O, osc, p = oracle_scene(w)
    frames_per_sample = {"c2": 4, "c3": 4, "c4": 1, "c5": 1}[args.config]
    poses = assets.sweep_poses(512) if w["sweep"] else None
    # the reference is single-threaded (pragmas commented out, scene.cpp:74,105); the oracle's OpenMP
    # variant over elements is used when it is actually faster on this host
    best_threads, best = 1, 0.0
    for th in sorted({1, threads}):
        O.oracle().orc_set_threads(th)
        t0 = time.perf_counter()
        osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=0)
        fps = 1.0 / (time.perf_counter() - t0)
        if fps > best:
            best, best_threads = fps, th
    O.oracle().orc_set_threads(best_threads)
    for k in range(args.warmup if args.config in ("c2", "c3") else 0):
        osc.simulate_frame(p, pose[:3], pose[3:], seed=1234, frame=k)
    t0 = time.perf_counter()
    segs = 0
    for k in range(args.steps):
        for f in range(frames_per_sample):
            i = k * frames_per_sample + f
            ps = poses[(i * 37) % 512] if poses is not None else pose
            segs += osc.simulate_frame(p, ps[:3], ps[3:], seed=1234, frame=100 + i)["tests"]
    dt = time.perf_counter() - t0
    fps = args.steps * frames_per_sample / dt
    sample = f"{frames_per_sample} frame(s) per step of the same workload, oracle port of the reference CPU path (the reference itself needs Bullet + OpenCV, not buildable here)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.config, w, world, "host CPU", None),
        "ray_segments_per_s": segs / dt,
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": best_threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def stage_counters(config: str = "c2") -> dict | None:
    """warp / thread instructions per stage and frame from the kept ncu counter file (profiles/stage_counters.json, written by
    scripts/collect_stage_counters.py): instruction counts per frame do not depend on the run, so the live stage times turn
    them into issue-slot utilisation"""
    for p in (ROOT / "profiles" / f"stage_counters_{config}.json", ROOT / "profiles" / "stage_counters.json"):
        try:
            d = json.loads(p.read_text())
            if d.get("config") == config:
                d["file"] = "profiles/" + p.name
                return d
        except Exception:
            pass
    return None


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mcray_tracing_b200 import api, assets, sweep

    if rank == 0:
        assets.ensure_all()
    if world > 1:
        dist.barrier()
    w = workload(args.config, args.frames_per_step)
    strong = args.config == "c3"
    if strong:
        if 512 % world:
            raise SystemExit("bench.py: --config c3 shards 512 poses: --gpus must divide 512")
        w["F"] = 512 // world
    F = w["F"]
    params = api.default_params(**w["params"])
    sim = api.Simulator(w["scene"], params, device=local_rank)
    sim.set_option("max_batch_poses", max(F, 1))
    for ov in args.option:
        oname, oval = ov.split("=")
        sim.set_option(oname, int(oval))
    # the background-built SAH tree (mcrt.h "bvh_optimise") is adopted before anything is timed: no tree swap inside a timed region
    t_wait = time.perf_counter()
    sim.set_option("bvh_wait", 1)
    bvh_state = {"optimised": int(sim.get_info().bvh_optimised), "wait_s": round(time.perf_counter() - t_wait, 3)}
    rows, cols = sim.rows, sim.cols
    total = F * world                                                     # frames per step, whole job
    base_pose = w["pose"] if w["pose"] is not None else sim.start_pose
    use_sweep = w["sweep"] or (world > 1 and args.config == "c2")
    interleave = world > 1 and not args.contiguous
    if use_sweep:
        allp = assets.sweep_poses(512 if strong else total)
        idx = sweep.shard_indices(len(allp), world, rank, interleave)
        poses = allp[idx]
        my_first = int(idx[0]) if len(idx) else 0
    else:
        allp = np.repeat(base_pose[None, :], total, axis=0)
        idx = sweep.shard_indices(total, world, rank, interleave)
        poses, my_first = allp[idx], int(idx[0])
    stride = world if interleave else 1
    sim.set_option("frame_stride", stride)
    seed = 1234
    st = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)
    # N > 1: every rank DEPOSITS its finished RF lines straight into rank 0's double-buffered receive buffer over NVLink
    # (sweep.PeerDeposit: CUDA-IPC peer copies + a 4-byte NCCL all-reduce as the completion signal; with the round-robin deal
    # one strided copy puts the frames in global pose order); --gather nccl selects the grouped ncclSend/ncclRecv gather.
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
        if rank == 0 and not interleave:
            outs = [r[:F] for r in recvs]                                 # rank 0 simulates in place
        else:
            outs = [torch.empty((F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2)]
    elif world > 1 and rank == 0:
        recvs = [torch.empty((world * F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2)]
        outs = [r[:F] for r in recvs]
    else:
        outs = [torch.empty((F, cols, rows), dtype=torch.float32, device=dev) for _ in range(2 if world > 1 else 1)]
    out = outs[0]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    gather_done = [torch.cuda.Event() for _ in range(2)]
    pending = [False, False]

    def first_frame(k: int) -> int:
        return k * total + my_first

    # round-robin deal on the collecting rank: its frames go STRAIGHT into their interleaved slots of the receive buffer (the post kernel
    # writes them there: option rf_out_frame_stride), no local copy at all
    direct_slot = peer is not None and interleave and rank == 0 and recvs is not None

    def compute(k: int, i: int):
        """simulate this rank's poses of step k into outs[i % 2] on stream st"""
        b = i % len(outs)
        if pending[b]:                                   # the gather that last read this buffer must be done
            st.wait_event(gather_done[b]); pending[b] = False
        if direct_slot:
            sim.set_option("rf_out_frame_stride", world)
            sim.simulate_device(poses, recvs[b].data_ptr() + rank * cols * rows * 4, seed=seed, first_frame=first_frame(k), stream=st.cuda_stream, sync=False)
            sim.set_option("rf_out_frame_stride", 1)
            return
        sim.simulate_device(poses, outs[b].data_ptr(), seed=seed, first_frame=first_frame(k), stream=st.cuda_stream, sync=False)

    def gather(i: int, after: "torch.cuda.Event"):
        """bring the finished RF lines of outs[i % 2] to rank 0 on the comm stream, not before `after`"""
        b = i % len(outs)
        comm.wait_event(after)
        if args.gather == "none":
            gather_done[b].record(comm)
        elif peer is not None:
            if not os.environ.get("MCRT_DIAG_SKIP_DEPOSIT") and not direct_slot:   # (env: diagnostics of the gather's cost, profiles/r02ai_*)
                peer.deposit(b, outs[b], comm, interleave=interleave)
            if not os.environ.get("MCRT_DIAG_SKIP_COMMIT"):
                peer.commit(comm)
            gather_done[b].record(comm)
        else:
            with torch.cuda.stream(comm):
                sweep.gather_lines(outs[b], [F] * world, dst=0, out=recvs[b] if recvs is not None else None, in_place=recvs is not None)
                gather_done[b].record(comm)
        pending[b] = True

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    done_ev = torch.cuda.Event()
    n_warm = max(args.warmup, 3)
    for k in range(n_warm):
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
    value = total * args.steps / (total_ms * 1e-3)
    per_rank_ms = [my_ms / args.steps]
    if world > 1:
        allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([my_ms / args.steps], dtype=torch.float64, device=dev))
        per_rank_ms = [float(t.item()) for t in allms]

    # ---- N > 1: is the gathered result the 1-GPU result?  rank 0 re-simulates a sample of the last step's frames (every
    # rank's share) on its own GPU and compares them, bit for bit, with what landed in its receive buffer ---------------------
    multi_check = None
    if world > 1 and rank == 0 and args.gather == "none":
        multi_check = {"invalid_for_scaling": "diagnostic run without any exchange (--gather none)"}
    elif world > 1 and rank == 0:
        k_last = 1000 + args.steps - 1
        got = recvs[(args.steps - 1) % 2]
        order_is_global = (peer is not None and interleave) or not interleave   # NCCL gather of a round-robin deal stays in rank-block order
        sample = sorted({int(x) for x in np.linspace(0, total - 1, 4 * world)})
        sim.set_option("frame_stride", 1)
        ref1 = torch.empty((1, cols, rows), dtype=torch.float32, device=dev)
        same = True
        hsh = hashlib.sha256()
        for g in sample:
            sim.simulate_device(allp[g:g + 1], ref1.data_ptr(), seed=seed, first_frame=k_last * total + g)
            slot = g if order_is_global else (g % world) * F + g // world
            same = same and bool(torch.equal(ref1[0], got[slot]))
            hsh.update(got[slot].cpu().numpy().tobytes())
        sim.set_option("frame_stride", stride)
        multi_check = {"n_gpu_bit_identical": same, "frames_checked": len(sample), "of_frames": total,
                       "sha256_of_checked_frames": hsh.hexdigest(), "gathered_sum": float(got.double().sum().item()),
                       "how": "rank 0 re-simulated these frames of the last timed step on one GPU and compared them with its gathered buffer (torch.equal)"}

    # ---- e2e: host buffers, copies inside the timed region -----------------------------------------
    host_out = torch.empty((F, cols, rows), dtype=torch.float32, pin_memory=True)
    host_np = host_out.numpy()
    for k in range(2):
        sim.simulate(poses, seed=seed, first_frame=first_frame(k), rf_out=host_np)
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sim.simulate(poses, seed=seed, first_frame=first_frame(2000 + k), rf_out=host_np)
        checksum = float(host_np[0, cols // 2, rows // 2])                 # the step's result is read on the host
    e2e_sync_s = time.perf_counter() - t0
    # the same through the streaming driver (stream.FrameStreamer, public API): every step still uploads its
    # poses from host memory and lands its RF frames in pinned host memory, but the PCIe copy of step k
    # overlaps the simulation of step k+1 (separate copy stream, ring of 3 pinned buffers)
    from mcray_tracing_b200 import stream as mstream
    e2e_sub = max(1, min(int(args.e2e_sub_batches), F))
    fs = mstream.FrameStreamer(sim, depth=3, frames_per_submit=F, seed=seed, sub_batches=e2e_sub, frame_stride=stride)

    def stream_steps(n_steps: int, k0: int) -> float:
        submitted, chk = 0, 0.0
        t0 = time.perf_counter()
        while submitted < n_steps or fs.pending():
            while submitted < n_steps and fs._free:
                fs._frame = first_frame(k0 + submitted)
                fs.submit(poses)
                submitted += 1
            _, rf_host, _ = fs.get()
            chk += float(rf_host[0, cols // 2, rows // 2])               # the step's result is read on the host
        return time.perf_counter() - t0

    stream_steps(3, 4000)
    sync_all()
    e2e_s = stream_steps(args.steps, 5000)
    # the ceiling of that path: the same bytes, device -> the same pinned ring, nothing else (all ranks at once)
    d2h_bytes = F * cols * rows * 4
    sync_all()
    cp = torch.cuda.Stream(device=dev)
    ring = [s["rf_host"] for s in fs._slots]
    src = fs._slots[0]["rf_dev"]
    with torch.cuda.stream(cp):
        for k in range(2):
            ring[k % 3].copy_(src, non_blocking=True)
    cp.synchronize()
    sync_all()
    t0 = time.perf_counter()
    with torch.cuda.stream(cp):
        for k in range(args.steps):
            ring[k % 3].copy_(src, non_blocking=True)
    cp.synchronize()
    d2h_s = time.perf_counter() - t0
    # declared alternative payload: the 8-bit scan-converted B-mode image (mcrt_bmode: TGC, log compression, scan conversion,
    # quantisation; 200 KB per frame instead of 476 KB of float RF lines)
    bmode_s = None
    try:
        import ctypes as C
        L = api.lib()
        bp = api.BmodeParams(0.0, 0.0, 60.0, 0.0)
        img8 = torch.empty((F, sim.info.scan_rows, sim.info.scan_cols), dtype=torch.uint8, pin_memory=True)
        def bmode_step(k):
            sim.simulate_device(poses, src.data_ptr(), seed=seed, first_frame=first_frame(k))
            api._check(L.mcrt_bmode(sim.h, C.c_void_p(src.data_ptr()), F, C.byref(bp), None, C.c_void_p(img8.data_ptr())))
            return int(img8[0, 10, 10])
        bmode_step(6000)
        sync_all()
        t0 = time.perf_counter()
        for k in range(args.steps):
            bmode_step(6001 + k)
        bmode_s = time.perf_counter() - t0
    except Exception as e:                                                  # the alternative payload is optional
        print(f"bench.py: b-mode payload variant skipped ({e})", file=sys.stderr)
    # N > 1: the GATHERED sweep on rank 0's host (what a single consumer process sees): simulate, deposit, then rank 0 copies
    # all N x F frames to pinned host memory
    gathered_s = None
    if world > 1 and peer is not None:
        ghost = torch.empty((world * F, cols, rows), dtype=torch.float32, pin_memory=True) if rank == 0 else None
        def gathered_step(k, b):
            sim.simulate_device(poses, outs[b].data_ptr(), seed=seed, first_frame=first_frame(k), stream=st.cuda_stream, sync=False)
            done_ev.record(st)
            comm.wait_event(done_ev)
            peer.deposit(b, outs[b], comm, interleave=interleave)
            peer.commit(comm)
            if rank == 0:
                with torch.cuda.stream(comm):
                    ghost.copy_(recvs[b], non_blocking=True)
            comm.synchronize()
        gathered_step(7000, 0)
        sync_all()
        t0 = time.perf_counter()
        for k in range(args.steps):
            gathered_step(7001 + k, k % 2)
        sync_all()
        gathered_s = time.perf_counter() - t0
    tt = [e2e_s, e2e_sync_s, d2h_s, bmode_s if bmode_s is not None else 0.0, gathered_s if gathered_s is not None else 0.0]
    e2e_t = torch.tensor(tt, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = total * args.steps / float(e2e_t[0].item())
    e2e_sync_value = total * args.steps / float(e2e_t[1].item())
    d2h_gbs_per_gpu = d2h_bytes * args.steps / float(e2e_t[2].item()) / 1e9
    e2e_gbs_per_gpu = d2h_bytes * args.steps / float(e2e_t[0].item()) / 1e9

    # ---- latency mode (one frame per call) and per-stage kernel times (rank 0 only) ---------------
    extra = {}
    roofline = None
    cpu_baseline = None
    if rank == 0:
        sim.set_option("frame_stride", 1)
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
        # the headline's neighbour (c2, one GPU): the same call size on the POSES OF THE PROBE SWEEP -- the workload every rank runs at
        # --gpus N > 1 -- so that the N-GPU values have a like-for-like single-GPU figure beside them
        if world == 1 and args.config == "c2":
            sp = assets.sweep_poses(F)
            sw = []
            for k in range(2 + 5):
                flush.fill_(0.0)
                torch.cuda.synchronize(dev)
                sim.simulate_device(sp, out.data_ptr(), seed=seed, first_frame=3000 + k * F)
                if k >= 2:
                    sw.append(sim.stats().ms_total)
            extra["sweep_workload_1gpu"] = {"value": F / float(np.mean(sw)) * 1e3, "unit": UNIT, "ms_per_step": float(np.mean(sw)), "frames_per_step": F,
                                            "poses": f"the first {F} poses of the ircad11 probe sweep (what each rank simulates at --gpus N > 1)"}
        # per-stage device times: separate pass, stage events between the kernels (no CUDA graph)
        sim.set_option("profile_stages", 1)
        tr, ac, po, tot, msteps = [], [], [], [], []
        for k in range(2 + 5):
            flush.fill_(0.0)
            torch.cuda.synchronize(dev)
            sim.simulate_device(poses, out.data_ptr(), seed=seed, first_frame=(3000 + k) * total + my_first)
            s = sim.stats()
            if k >= 2:
                tr.append(s.ms_trace); ac.append(s.ms_accumulate); po.append(s.ms_post); tot.append(s.ms_total); msteps.append(s.march_steps)
        # real closest-hit traversals: with the bounce-0 de-duplication a bounce-0 ray is traced once per (pose, element)
        sim.set_option("count_traversal", 1)
        sim.simulate_device(poses, out.data_ptr(), seed=seed, first_frame=3100 * total + my_first)
        sct = sim.stats()
        sim.set_option("count_traversal", 0)
        sim.set_option("profile_stages", 0)
        ms_tr, ms_acc, ms_po = float(np.mean(tr)), float(np.mean(ac)), float(np.mean(po))
        peak, peak_src = measured_hbm_peak()
        sm_hz = (clock_info.get("sm_mhz") or 1965.0) * 1e6
        issue_peak = sim.info.sm_count * 4 * sm_hz                            # warp instructions per second the SMs can issue
        S = w["params"]["samples"]
        dedup = S >= 4 and F * cols >= 4096
        queries = seg_per_step - (F * cols * (S - 1) if dedup else 0)
        sc = stage_counters(args.config)
        stages = []
        for name, ms in (("trace", ms_tr), ("accumulate", ms_acc), ("post", ms_po)):
            e = {"stage": name, "ms": ms, "share_of_step": ms / (ms_tr + ms_acc + ms_po)}
            cnt = (sc or {}).get("stages", {}).get(name) if sc and sc.get("config") == args.config else None
            if cnt:
                scale = F / float(sc["frames_per_launch"])
                wi, ti = cnt["warp_inst"] * scale, cnt["thread_inst"] * scale
                e.update({"warp_inst": wi, "issue_frac": wi / (ms * 1e-3) / issue_peak, "lanes_per_inst": ti / wi,
                          "counters": sc.get("file", "profiles/stage_counters.json") + " (ncu smsp__inst_executed.sum / smsp__thread_inst_executed.sum of one step)"})
                if name == "trace":
                    e["thread_inst_per_query"] = ti / queries
                if name == "accumulate":
                    e["thread_inst_per_march_step"] = ti / float(np.mean(msteps))
                if name == "post":
                    e["thread_inst_per_pixel"] = ti / (F * cols * rows)
            stages.append(e)
        stages.sort(key=lambda e: -e["ms"])
        # HBM roofline of the memory-streaming kernel of this configuration: accumulate = 8 B per march step
        # + 4 B per RF sample; c5: the post stage = 8 B per pixel when fused (read once, write once)
        alg_acc = 8.0 * float(np.mean(msteps)) + 4.0 * F * cols * rows
        alg_post = 8.0 * F * cols * rows
        traffic = None
        tf = ROOT / "profiles" / "accumulate_traffic.json"
        if tf.exists() and args.config in ("c2", "c3"):
            try:
                tj = json.loads(tf.read_text())
                traffic = int(tj["dram_bytes_per_launch"] * F / tj.get("frames_per_launch", F))
            except Exception:
                traffic = None
        if args.config == "c5":
            ka, kl = w["params"]["psf_axial"], w["params"]["psf_lateral"]
            flops = 2.0 * (ka + kl) * F * cols * rows
            fp32_peak = sim.info.sm_count * 128 * sm_hz                       # un-contracted: one multiply or add per lane and clock
            t_roof = max(alg_post / (peak * 1e9), flops / fp32_peak)
            roofline = {"kernel": "post stage (axial + lateral PSF + envelope)", "bound": "hbm", "achieved": alg_post / (ms_po * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg_post / (ms_po * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg_post, "kernel_ms": ms_po,
                        "fp32_flops_per_launch": flops, "fp32_instr_peak_per_s": fp32_peak,
                        "frac_of_max_bytes_flops_roof": t_roof / (ms_po * 1e-3),
                        "note": "63 x 31 taps without FMA contraction (bit-exactness) = 188 instructions per pixel: fp32-issue-bound, as SURVEY 8(d) predicts; "
                                "frac_of_max_bytes_flops_roof = max(bytes / HBM peak, flops / fp32 issue peak) / measured time"}
        else:
            roofline = {"kernel": "k_accumulate_win (echo accumulation + scatterer-volume gather + sample reduction, one kernel)", "bound": "hbm",
                        "achieved": alg_acc / (ms_acc * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_acc / (ms_acc * 1e-3) / 1e9 / peak,
                        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_acc, "kernel_ms": ms_acc,
                        "note": "the HBM-streaming kernel of the step; it is issue-bound (volume L2-resident), see stages[].issue_frac"}
        roofline["stage_ms"] = {"trace": ms_tr, "accumulate": ms_acc, "post": ms_po, "total": float(np.mean(tot))}
        roofline["stages"] = stages                                            # time-dominant stage first
        extra["roofline_trace"] = {"kernel": "k_first_hit + k_bounce x max_depth + k_compact (BVH traversal + boundary physics + compaction)",
                                   "bound": "sm_issue", "stage_ms": ms_tr, "closest_hit_queries_per_step": queries,
                                   "closest_hit_queries_per_s": queries / (ms_tr * 1e-3), "reference_equivalent_segments_per_s": seg_per_step / (ms_tr * 1e-3),
                                   "bvh_node_visits_per_query": sct.bvh_node_visits / max(queries, 1), "triangle_tests_per_query": sct.bvh_triangle_tests / max(queries, 1),
                                   "issue_frac": next((e.get("issue_frac") for e in stages if e["stage"] == "trace"), None),
                                   "thread_inst_per_query": next((e.get("thread_inst_per_query") for e in stages if e["stage"] == "trace"), None),
                                   "note": "bound by SM instruction issue (ALU pipe) and divergence, not by bandwidth: issue_frac = warp instructions / "
                                           "(SMs x 4 schedulers x SM clock x stage time)"}
        if args.config == "c5":
            # the post kernels alone on a synthetic 1024 x 8192 image (exactly BASELINE's size)
            rng = np.random.default_rng(7)
            img = rng.standard_normal((1024, 8192)).astype(np.float32)
            ax, la = sim.psf_taps()
            t_post = []
            for k in range(4):
                t0 = time.perf_counter(); sim.postprocess(img, ax, la); t_post.append(time.perf_counter() - t0)
            extra["post_only_1024x8192"] = {"wall_ms_incl_h2d_d2h": float(np.min(t_post)) * 1e3,
                                            "note": "mcrt_postprocess on host buffers: dominated by the 2 x 33.5 MB PCIe copies; device time is in roofline.kernel_ms per frame"}
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            n = args.cpu_frames if args.config in ("c2", "c3") else 3
            fps1, sps1, stage1, n1 = oracle_frames_per_s(w, base_pose, n, 1)
            fpsN, spsN, stageN, nN = (fps1, sps1, stage1, n1) if threads == 1 else oracle_frames_per_s(w, base_pose, n, threads)
            best_fps, best_sps, cores = (fpsN, spsN, threads) if fpsN > fps1 else (fps1, sps1, 1)
            cpu_baseline = {"value": best_fps, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n1} / {nN} frames of the same workload (single thread / all threads, each bounded to ~25 s); single thread (as the "
                                      f"reference ships, scene.cpp:74) = {fps1:.3f} frames/s, OpenMP over elements on {threads} threads = {fpsN:.3f} frames/s; "
                                      "oracle port (the reference needs Bullet + OpenCV)",
                            "ray_segments_per_s": best_sps,
                            "single_thread": {"frames_per_s": fps1, "ray_segments_per_s": sps1, "stage_ms_per_frame": stage1},
                            "all_threads": {"threads": threads, "frames_per_s": fpsN, "ray_segments_per_s": spsN, "stage_ms_per_frame": stageN},
                            "stage_order": ["cast_rays", "accumulate", "convolve+envelope", "scan_convert"]}
    if rank == 0:
        par = "single GPU" if world == 1 else (f"poses dealt round-robin x{world}" if interleave else f"contiguous pose blocks x{world}")
        gat = None if world == 1 else ("p2p deposit over NVLink (CUDA IPC, one strided copy per rank and step) + 4-byte NCCL all-reduce" if peer is not None else "NCCL grouped send/recv")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(workload_config(args.config, w, world, par, gat),
                                                 bvh=("device LBVH at mcrt_create, replaced by the host binned-SAH tree built in the background (adopted before the "
                                                      f"first timed step; waited {bvh_state['wait_s']} s)" if bvh_state["optimised"] else "device LBVH"),
                                                 **({"options": args.option} if args.option else {})),
            "ray_segments_per_s": float(segs_all.item()) * args.steps / (total_ms * 1e-3),
            "segments_per_step": float(segs_all.item()), "march_steps_per_step_per_gpu": march_per_step,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * 24 + 16), "d2h_bytes_per_step": int(d2h_bytes),
                    "timing": "wall clock around K steps through stream.FrameStreamer (mcrt_simulate_async + pinned ring of 3 buffers): "
                              "per step poses host->device and RF frames device->pinned host, the copy of step k overlapping step k+1"
                              + (f"; every step is submitted as {e2e_sub} consecutive calls whose copies start as soon as each is computed" if e2e_sub > 1 else ""),
                    "sub_batches_per_step": e2e_sub,
                    "synchronous_call_value": e2e_sync_value,
                    "synchronous_call_timing": "wall clock around K blocking mcrt_simulate calls with pinned host buffers (no overlap)",
                    "roofline": {"bound": "pcie_d2h", "ceiling_gbs_per_gpu": d2h_gbs_per_gpu, "achieved_gbs_per_gpu": e2e_gbs_per_gpu,
                                 "frac": e2e_gbs_per_gpu / d2h_gbs_per_gpu, "ceiling_frames_per_s": total * args.steps / float(e2e_t[2].item()),
                                 "how": "the same bytes per step, device -> the same pinned ring with plain async copies, all ranks at once, max over ranks"},
                    "bmode8_payload": None if bmode_s is None else {"value": total * args.steps / float(e2e_t[3].item()), "unit": UNIT,
                                      "d2h_bytes_per_step": int(F * sim.info.scan_rows * sim.info.scan_cols),
                                      "what": "declared alternative payload: simulate + mcrt_bmode (TGC, log compression, scan conversion, 8-bit) per step, blocking calls"},
                    "gathered_on_rank0_host": None if gathered_s is None else {"value": total * args.steps / float(e2e_t[4].item()), "unit": UNIT,
                                      "d2h_bytes_per_step_rank0": int(world * d2h_bytes),
                                      "what": "every step: simulate, deposit on rank 0 over NVLink, rank 0 copies the whole gathered sweep to pinned host memory"},
                    "host_numa_binding": numa},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "wall_s_timed_region": t_wall,
            "ms_per_step_per_rank": per_rank_ms,
        }
        if multi_check is not None:
            line.update(multi_check)
        if "sweep_workload_1gpu" in extra:                                      # beside the headline, not among the extras
            items = list(line.items())
            i = [k for k, _ in items].index("ms_per_step") + 1
            line = dict(items[:i] + [("sweep_workload_1gpu", extra.pop("sweep_workload_1gpu"))] + items[i:])
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
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"], help="BASELINE.json configuration (default c2: the one the metric is quoted on)")
    ap.add_argument("--frames-per-step", type=int, default=None,
                    help="independent frames per C-ABI call and per GPU (defaults: c2 1024, c3 512 / N, c4 128, c5 32; "
                         "measured on one B200: 256 -> 99k, 512 -> 107k, 1024 -> 112k frames/s)")
    ap.add_argument("--cpu-frames", type=int, default=60, help="frames per CPU-baseline variant (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"],
                    help="N > 1: peer-memory deposit over NVLink (default), NCCL send/recv gather, or none (diagnostic: no exchange at all, "
                         "the line is marked invalid_for_scaling)")
    ap.add_argument("--e2e-sub-batches", type=int, default=2, help="e2e leg: calls per step of the streaming driver (stream.FrameStreamer sub_batches)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="development A/B: mcrt_set_option before the run (recorded in config.options)")
    ap.add_argument("--contiguous", action="store_true", help="N > 1: contiguous pose blocks per rank instead of the round-robin deal")
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
