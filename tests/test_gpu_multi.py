"""Multi-GPU (one process per GPU, NCCL) checks of both partitions of SURVEY 8(e) on real devices:
the pose-sharded sweep and the scanline-block split of one frame, each bit-identical to the 1-GPU result.
Skipped on a 1-GPU box (the same host logic runs on gloo / CPU in test_host_and_numerics.py)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from mcray_tracing_b200 import api, assets, sweep
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
d = assets.ensure_all()
scene = d["ircad11"] / "santi-liver-rough.scene"
sim = api.Simulator(scene, api.default_params(elements=96, samples=4), device=rank)
poses = assets.sweep_poses(37)                                   # ragged over 2 ranks
def block(p, first_frame, out):
    sim.simulate_device(p, out.data_ptr(), seed=21, first_frame=first_frame)
res = sweep.run_sweep(block, poses, (sim.cols, sim.rows), dev, seed_first_frame=300)
pose = sim.start_pose
def lines(first, n, out):
    sim.simulate_scanlines(pose, first, n, seed=21, frame=9, rf_ptr=out.data_ptr())
frame = sweep.run_frame_scanline_blocks(lines, sim.cols, sim.rows, dev)
# the same sweep through the peer-memory deposit (CUDA IPC copies over NVLink instead of the NCCL gather)
b, e = sweep.shard_bounds(len(poses), world, rank)
F = max(sweep.all_shard_sizes(len(poses), world))
peer = sweep.PeerDeposit(F, (sim.cols, sim.rows), dev, n_slots=2, dst=0)
comm = torch.cuda.Stream(device=dev)
local = torch.zeros((F, sim.cols, sim.rows), dtype=torch.float32, device=dev)
sim.simulate_device(poses[b:e], local.data_ptr(), seed=21, first_frame=300 + b)
torch.cuda.synchronize(dev)
for slot in (1, 0, 1):
    peer.deposit(slot, local, comm)
    peer.commit(comm)
comm.synchronize()
deposited = None
if rank == 0:
    t = peer.slot_tensor(1).cpu().numpy()
    sizes = sweep.all_shard_sizes(len(poses), world)
    deposited = np.concatenate([t[r * F: r * F + sizes[r]] for r in range(world)])
# round-robin deal (rank r: poses r, r + G, ...; option frame_stride = G), NCCL gather and strided peer deposit
sim.set_option("frame_stride", world)
res_il = sweep.run_sweep(block, poses, (sim.cols, sim.rows), dev, seed_first_frame=300, interleave=True)
idx = sweep.shard_indices(len(poses), world, rank, interleave=True)
local.zero_()
sim.simulate_device(poses[idx], local.data_ptr(), seed=21, first_frame=300 + rank)
sim.set_option("frame_stride", 1)
torch.cuda.synchronize(dev)
peer.deposit(0, local[:len(idx)], comm, interleave=True)
peer.commit(comm)
comm.synchronize()
deposited_il = peer.slot_tensor(0).cpu().numpy()[:len(poses)] if rank == 0 else None
peer.close()
if rank == 0:
    single = sim.simulate(poses, seed=21, first_frame=300)
    assert np.array_equal(res_il.cpu().numpy(), single), "round-robin sweep (NCCL gather) differs from the 1-GPU sweep"
    assert np.array_equal(deposited_il, single), "round-robin sweep (strided peer deposit) differs from the 1-GPU sweep"
    assert np.array_equal(deposited, single), "peer-deposited sweep differs from the 1-GPU sweep"
    assert np.array_equal(res.cpu().numpy(), single), "pose-sharded sweep differs from the 1-GPU sweep"
    whole = sim.simulate(pose[None, :], seed=21, first_frame=9)[0]
    assert np.array_equal(frame.cpu().numpy(), whole), "scanline-block frame differs from the 1-GPU frame"
    print("OK multi", world)
else:
    assert res is None and frame is None and res_il is None
sim.close()
dist.destroy_process_group()
"""


def test_sweep_and_scanline_blocks_two_gpus_nccl(tmp_path, built, assets_dirs):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker_multi.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    assert "OK multi 2" in outs[0][0]
