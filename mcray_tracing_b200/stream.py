"""Streaming driver: a stream of probe poses in, RF (and optionally scan-converted) frames out,
replacing the reference's interactive loop (inputmanager.cpp:61-122 moves the probe, main.cpp:92-152
re-simulates and rf_image::show() displays -- SURVEY.md section 8f item 1).

Frames are enqueued on a dedicated compute stream through mcrt_simulate_async; their results are
copied into a ring of PINNED host buffers on a second (copy) stream that waits on the frame's event, so
frame k+1 is being traced while frame k is still travelling over PCIe; `get()` only waits for the event
of the frame it returns.
The pipeline depth bounds the latency: a frame is at most `depth` submissions behind.
With `sub_batches` > 1 a submission is simulated as that many consecutive calls and every part starts its
PCIe copy as soon as it is computed, while the next part is being simulated: the copy that nothing can
overlap (the last one of a run) shrinks to one part.  Frames are pure functions of (pose, seed, frame
index), so the split does not change them.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

from . import api


class FrameStreamer:
    def __init__(self, sim: api.Simulator, depth: int = 3, frames_per_submit: int = 1, scan: bool = False, seed: int = 0,
                 device: int | None = None, sub_batches: int = 1, frame_stride: int = 1):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one frame in flight while one is read)")
        self.sim = sim
        self.n = int(frames_per_submit)
        if not 1 <= int(sub_batches) <= self.n:
            raise ValueError("sub_batches must be in [1, frames_per_submit]")
        # part boundaries (as even as possible); frame_stride = the context's "frame_stride" option (pose i of a call is frame first + i * stride)
        self._parts = [(k * self.n // int(sub_batches), (k + 1) * self.n // int(sub_batches)) for k in range(int(sub_batches))]
        self.frame_stride = int(frame_stride)
        self.scan = bool(scan)
        self.seed = int(seed)
        self.dev = torch.device("cuda", sim.info.device if device is None else device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        shape = sim.rf_shape(self.n)
        sshape = (self.n, sim.info.scan_rows, sim.info.scan_cols)
        self._slots = []
        for _ in range(depth):
            slot = dict(rf_dev=torch.empty(shape, dtype=torch.float32, device=self.dev),
                        rf_host=torch.empty(shape, dtype=torch.float32, pin_memory=True),
                        computed=torch.cuda.Event(), done=torch.cuda.Event())
            if self.scan:
                slot["scan_dev"] = torch.empty(sshape, dtype=torch.float32, device=self.dev)
                slot["scan_host"] = torch.empty(sshape, dtype=torch.float32, pin_memory=True)
            self._slots.append(slot)
        self._free = deque(range(depth))
        self._inflight: deque[tuple[int, int]] = deque()      # (ticket, slot)
        self._ticket = 0
        self._frame = 0

    def submit(self, poses) -> int:
        """Enqueue `frames_per_submit` poses; returns a ticket.  Blocks only if every slot is in flight
        (then the oldest result must be fetched with get() first)."""
        P = api.make_poses(poses)
        if len(P) != self.n:
            raise ValueError(f"expected {self.n} poses per submit")
        if not self._free:
            raise RuntimeError("pipeline full: call get() before submitting more frames")
        s = self._free.popleft()
        slot = self._slots[s]
        for a, b in self._parts:
            self.sim.simulate_device(P[a:b], slot["rf_dev"][a:].data_ptr(), seed=self.seed, first_frame=self._frame + a * self.frame_stride,
                                     scan_ptr=slot["scan_dev"][a:].data_ptr() if self.scan else None, stream=self.stream.cuda_stream, sync=False)
            slot["computed"].record(self.stream)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(slot["computed"])
                slot["rf_host"][a:b].copy_(slot["rf_dev"][a:b], non_blocking=True)
                if self.scan:
                    slot["scan_host"][a:b].copy_(slot["scan_dev"][a:b], non_blocking=True)
        with torch.cuda.stream(self.copy_stream):
            slot["done"].record(self.copy_stream)
        self._frame += self.n
        self._ticket += 1
        self._inflight.append((self._ticket, s))
        return self._ticket

    def pending(self) -> int:
        return len(self._inflight)

    def get(self):
        """Oldest submitted result: (ticket, rf [n, ...] numpy view of pinned memory, scan or None).
        The arrays stay valid until `depth - 1` further submits."""
        if not self._inflight:
            raise RuntimeError("nothing in flight")
        ticket, s = self._inflight.popleft()
        slot = self._slots[s]
        slot["done"].synchronize()
        self._free.append(s)
        return ticket, slot["rf_host"].numpy(), (slot["scan_host"].numpy() if self.scan else None)

    def run(self, pose_iter, on_frame):
        """Convenience loop: keeps the pipeline full, calls on_frame(ticket, rf, scan) in order."""
        it = iter(pose_iter)
        exhausted = False
        while True:
            while not exhausted and self._free:
                try:
                    self.submit(next(it))
                except StopIteration:
                    exhausted = True
            if not self._inflight:
                break
            on_frame(*self.get())


def sweep_pose_stream(poses: np.ndarray, frames_per_submit: int = 1):
    for i in range(0, len(poses) - frames_per_submit + 1, frames_per_submit):
        yield poses[i:i + frames_per_submit]
