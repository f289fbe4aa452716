"""Probe-sweep sharding across the GPUs of one box: one process per GPU (torch.distributed), poses
partitioned into contiguous blocks, every rank runs the whole per-frame chain for its block with
no data-path communication, then ONE collective gathers the finished RF lines on rank 0 (NCCL over
NVLink on GPUs; gloo in the CPU unit tests of this host logic).

Frames are pure functions of (scene, pose, seed, global frame index) -- the Philox counter carries
the *global* pose index -- so the gathered result is bit-identical to a 1-GPU run of the same sweep.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of rank `rank`; the first (n_items % world_size) ranks get one extra."""
    if world_size < 1 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_shard_sizes(n_items: int, world_size: int) -> list[int]:
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0] for r in range(world_size)]


def shard_indices(n_items: int, world_size: int, rank: int, interleave: bool = False) -> np.ndarray:
    """Global indices of the items of rank `rank`.  Contiguous blocks (shard_bounds), or dealt out round-robin (rank r takes
    r, r + G, r + 2G, ...): the cost of a frame varies along a freehand sweep, so contiguous blocks make the step as slow as
    the most expensive block (6 % at 8 GPUs, SCALE_r01), while the round-robin deal gives every rank a uniform sample."""
    if interleave:
        if world_size < 1 or not (0 <= rank < world_size) or n_items < 0:
            raise ValueError("bad shard arguments")
        return np.arange(rank, n_items, world_size, dtype=np.int64)
    b, e = shard_bounds(n_items, world_size, rank)
    return np.arange(b, e, dtype=np.int64)


def interleaved_sizes(n_items: int, world_size: int) -> list[int]:
    return [len(range(r, n_items, world_size)) for r in range(world_size)]


def gather_lines(local: torch.Tensor, sizes: list[int], group=None, dst: int = 0, out: torch.Tensor | None = None,
                 in_place: bool = False) -> torch.Tensor | None:
    """Gather per-rank blocks [sizes[r], ...] on `dst` into [sum(sizes), ...] with ONE collective: a group of point-to-point
    transfers to `dst` (what ncclGather is: grouped ncclSend / ncclRecv), so nobody but `dst` receives anything.
    Ragged blocks are padded to the largest block.
    `out` (rank `dst` only): a preallocated [world * max(sizes), ...] receive buffer, so a steady-state loop that runs the
    gather on its own stream allocates nothing.  `in_place`: on `dst`, `local` already IS out[dst * max : dst * max +
    sizes[dst]] (the simulation wrote straight into the receive buffer), so the self-copy is skipped."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    max_n = max(sizes)
    tail = tuple(local.shape[1:])
    if rank == dst:
        if out is None:
            out = local.new_empty((world * max_n,) + tail)
        elif tuple(out.shape) != (world * max_n,) + tail or not out.is_contiguous():
            raise ValueError("gather_lines: out must be a contiguous [world * max(sizes), ...] tensor")
        blocks = out.view((world, max_n) + tail)
        ops = [dist.P2POp(dist.irecv, blocks[r], r, group) for r in range(world) if r != dst]
        if not in_place:
            blocks[dst, : local.shape[0]].copy_(local)
        elif local.data_ptr() != blocks[dst].data_ptr():
            raise ValueError("gather_lines: in_place needs local to alias its block of out")
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if all(s == max_n for s in sizes):
            return out
        return torch.cat([out[r * max_n: r * max_n + sizes[r]] for r in range(world)], dim=0)
    if local.shape[0] != max_n:
        padded = local.new_zeros((max_n,) + tail)
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, padded, dst, group)]):
        w.wait()
    return None


def run_sweep(simulate_block: Callable[[np.ndarray, int, torch.Tensor], None], poses: np.ndarray, line_shape: tuple[int, ...],
              device: torch.device, seed_first_frame: int = 0, group=None, dst: int = 0, interleave: bool = False) -> torch.Tensor | None:
    """Shard `poses` ([n, 6]) over the ranks of `group`; `simulate_block(block_poses, first_frame, out)`
    must fill `out` ([len(block), *line_shape], on `device`) -- on a GPU that is
    api.Simulator.simulate_device(block, out.data_ptr(), first_frame=...).  Returns the gathered
    [n, *line_shape] tensor on rank `dst`, None elsewhere.
    interleave: rank r takes poses r, r + G, ... (see shard_indices); the frame of local pose i is then
    first_frame + i * G, i.e. the simulator must run with option frame_stride = G (simulate_block is called with
    first_frame = seed_first_frame + r)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(poses)
    if interleave and world > 1:
        idx = shard_indices(n, world, rank, True)
        local = torch.empty((len(idx),) + tuple(line_shape), dtype=torch.float32, device=device)
        if len(idx):
            simulate_block(poses[idx], seed_first_frame + rank, local)
        sizes = interleaved_sizes(n, world)
        got = gather_lines(local, sizes, group=group, dst=dst)
        if got is None:
            return None
        out = got.new_empty((n,) + tuple(line_shape))
        off = 0
        for r in range(world):                       # rank-block order -> global pose order
            out[r::world] = got[off:off + sizes[r]]
            off += sizes[r]
        return out
    b, e = shard_bounds(n, world, rank)
    local = torch.empty((e - b,) + tuple(line_shape), dtype=torch.float32, device=device)
    if e > b:
        simulate_block(poses[b:e], seed_first_frame + b, local)
    if world == 1:
        return local
    return gather_lines(local, all_shard_sizes(n, world), group=group, dst=dst)


def run_frame_scanline_blocks(simulate_block: Callable[[int, int, torch.Tensor], None], n_elements: int, rows: int,
                              device: torch.device, group=None, dst: int = 0) -> torch.Tensor | None:
    """ONE frame partitioned by scanline blocks (the single-pose latency partition): rank r simulates the
    contiguous scanline block `shard_bounds(n_elements, world, r)`; `simulate_block(first_element, n, out)`
    must fill `out` ([n, rows], on `device`) -- on a GPU that is
    api.Simulator.simulate_scanlines(pose, first_element, n, rf_ptr=out.data_ptr()), which re-traces the
    block's right-hand PSF halo locally (no halo exchange).  The same single gather as the pose sweep then
    assembles [n_elements, rows] on rank `dst` (None elsewhere); bit-identical to the unpartitioned frame."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = shard_bounds(n_elements, world, rank)
    local = torch.empty((e - b, rows), dtype=torch.float32, device=device)
    if e > b:
        simulate_block(b, e - b, local)
    if world == 1:
        return local
    return gather_lines(local, all_shard_sizes(n_elements, world), group=group, dst=dst)


class PeerDeposit:
    """The gather of finished RF lines without a data collective: rank `dst` owns `n_slots` receive buffers of
    [world * block_frames, *line_shape] float32 in its HBM and exports them over CUDA IPC; every other rank maps them and
    DEPOSITS its block with one peer-to-peer copy over NVLink (copy engines: no SM is spent, nobody but `dst` receives
    anything -- with grouped ncclSend/ncclRecv the collector's inbound rate capped 8-GPU runs at 122 GB/s).  A 4-byte
    all-reduce after the copies tells `dst` that the slot is complete.  Use: `deposit(slot, local, stream)` on every rank
    (rank `dst` may instead simulate straight into `local_block(slot)`), then `commit(stream)`; `slot_tensor(slot)` on `dst`."""

    def __init__(self, block_frames: int, line_shape: tuple[int, ...], device: torch.device, n_slots: int = 2, group=None, dst: int = 0):
        from . import api
        self.api, self.group, self.dst, self.device = api, group, dst, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.block_shape = (block_frames,) + tuple(line_shape)
        self.block_bytes = 4 * int(np.prod(self.block_shape))
        self.slot_bytes = self.block_bytes * self.world
        self.n_slots = n_slots
        self.base: list[int] = []
        handles = [None] * n_slots
        if self.rank == dst:
            try:
                for s in range(n_slots):
                    p = api.device_alloc(self.dev_index, self.slot_bytes)
                    self.base.append(p)
                    handles[s] = api.ipc_export(self.dev_index, p)
            except Exception as e:                        # the broadcast below must still happen, or the peers hang
                handles = ["error: %s" % e] * n_slots
        dist.broadcast_object_list(handles, src=dst, group=group)
        if isinstance(handles[0], str):
            raise RuntimeError("PeerDeposit: the collecting rank could not export its buffers (%s)" % handles[0])
        if self.rank != dst:
            self.base = [api.ipc_open(self.dev_index, h) for h in handles]
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)

    def block_ptr(self, slot: int, rank: int | None = None) -> int:
        return self.base[slot] + (self.rank if rank is None else rank) * self.block_bytes

    def deposit(self, slot: int, local: torch.Tensor, stream: torch.cuda.Stream, interleave: bool = False):
        """enqueue the copy of this rank's block into the collector's slot on `stream`.
        interleave: this rank holds the poses rank, rank + G, ... of the step (shard_indices): frame i goes to position
        i * G + rank of the slot -- one strided (2-D) copy, so the slot ends up in global pose order."""
        nbytes = local.numel() * local.element_size()
        if nbytes > self.block_bytes:
            raise ValueError("PeerDeposit.deposit: block larger than the slot")
        if interleave:
            frame_bytes = self.block_bytes // self.block_shape[0]
            self.api.copy2d_async(self.dev_index, self.base[slot] + self.rank * frame_bytes, self.world * frame_bytes, local.data_ptr(),
                                  frame_bytes, frame_bytes, local.shape[0], stream.cuda_stream)
            return
        if local.data_ptr() == self.block_ptr(slot):
            return                                        # already simulated in place (the collector itself)
        self.api.copy_async(self.dev_index, self.block_ptr(slot), local.data_ptr(), nbytes, stream.cuda_stream)

    def commit(self, stream: torch.cuda.Stream):
        """stream-ordered completion signal: when it has run on `dst`, every rank's deposit into the slot has landed"""
        with torch.cuda.stream(stream):
            dist.all_reduce(self._flag, group=self.group)

    def slot_tensor(self, slot: int) -> torch.Tensor:
        """the collector's view of a slot: [world * block_frames, *line_shape] (rank `dst` only)"""
        if self.rank != self.dst:
            raise RuntimeError("slot_tensor: only the collecting rank owns the buffers")

        class _Raw:
            pass
        raw = _Raw()
        shape = (self.world * self.block_shape[0],) + self.block_shape[1:]
        raw.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (self.base[slot], False), "version": 3, "strides": None}
        return torch.as_tensor(raw, device=self.device)

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self.base:
            if self.rank == self.dst:
                self.api.device_free(self.dev_index, p)
            else:
                self.api.ipc_close(self.dev_index, p)
        self.base = []
