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
              device: torch.device, seed_first_frame: int = 0, group=None, dst: int = 0) -> torch.Tensor | None:
    """Shard `poses` ([n, 6]) over the ranks of `group`; `simulate_block(block_poses, first_frame, out)`
    must fill `out` ([len(block), *line_shape], on `device`) -- on a GPU that is
    api.Simulator.simulate_device(block, out.data_ptr(), first_frame=...).  Returns the gathered
    [n, *line_shape] tensor on rank `dst`, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(poses)
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
