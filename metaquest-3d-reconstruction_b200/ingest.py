"""Streaming ingest of raw Quest depth frames: `.raw` files -> bounded ring of pinned host chunks
-> two device buffers, so that file reads, host-to-device copies and the K1/K2/K3 kernels of consecutive chunks
overlap and neither host nor device memory grows with the length of the capture.

The reference reads every frame several times through NumPy (`DepthDataIO.load_depth_map`,
dataio/depth_data_io.py:33-53: once in `build_depth_dataset` :166-261 for `is_depth_map_valid` :80-85, once per
frame in `integrate()`, 2R+1 times per frame in the confidence stage).  Here a side is read once per pass:

    reader threads --(file reads straight into a free ring slot)--> ready queue
    consumer: slot --cudaMemcpyAsync on a copy stream--> device buffer (c % 2) --event--> K1 (validity fused) -> ...

`RawDepthStreamer` is an iterator of (f0, f1, frames, present): `frames` is the chunk on the device (float32
[n,H,W]; valid until the next-but-one iteration) and `present` flags the files that existed.  With `device=None` the
chunks stay on the host (used by the CPU tests of the ring logic)."""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Optional, Sequence

import numpy as np


class RawDepthStreamer:
    """Ordered chunks of `<ts>.raw` frames through a bounded ring.

    path_of(i) -> path of frame i; n frames of `height` x `width` little-endian float32.  At most `slots` chunks of
    `chunk_frames` frames exist on the host at any time (pinned when a CUDA device is given) and two on the device.
    A missing file leaves a zero frame flagged `present == False` (the reference's `load_depth_map` returns None for
    it and the frame is skipped); a file of the wrong size raises `RuntimeError` in the consumer."""

    def __init__(self, path_of: Callable[[int], os.PathLike], n: int, height: int, width: int, chunk_frames: int = 64,
                 slots: int = 3, device=None, workers: Optional[int] = None):
        if slots < 2:
            raise ValueError("the ring needs at least two slots")
        self.path_of, self.n, self.H, self.W = path_of, int(n), int(height), int(width)
        self.chunk = max(1, int(chunk_frames))
        self.n_chunks = (self.n + self.chunk - 1) // self.chunk
        self.device = device
        self.workers = workers or min(16, os.cpu_count() or 1)
        self.slots = []
        self._torch = None
        if device is not None:
            import torch
            self._torch = torch
            self.slots = [torch.empty((self.chunk, self.H, self.W), dtype=torch.float32, pin_memory=True) for _ in range(slots)]
            self._np_slots = [s.numpy() for s in self.slots]
            self.dev_bufs = [torch.empty((self.chunk, self.H, self.W), dtype=torch.float32, device=device) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream(device=device)
        else:
            self._np_slots = [np.empty((self.chunk, self.H, self.W), np.float32) for _ in range(slots)]
        self.host_bytes = len(self._np_slots) * self.chunk * self.H * self.W * 4
        self._free: "queue.Queue[int]" = queue.Queue()
        for s in range(len(self._np_slots)):
            self._free.put(s)
        self._ready: "queue.Queue[tuple]" = queue.Queue()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self.max_slots_in_use = 0
        self.host_view = None
        self._in_use = 0
        self._lock = threading.Lock()

    # -- producer -------------------------------------------------------------------------------------------
    def _read(self, args):
        slot, row, i = args
        p = self.path_of(i)
        buf = self._np_slots[slot][row]
        if not os.path.exists(p):
            buf[...] = 0.0
            return False
        a = np.fromfile(p, dtype="<f4")
        if a.size != self.H * self.W:
            raise RuntimeError(f"{p}: expected {self.H * self.W} float32 values, found {a.size}")
        buf[...] = a.reshape(self.H, self.W)
        return True

    def _produce(self):
        try:
            with ThreadPoolExecutor(max_workers=self.workers) as pool:     # file reads release the GIL
                for c in range(self.n_chunks):
                    slot = None
                    while slot is None:
                        if self._stop.is_set():
                            return
                        try:
                            slot = self._free.get(timeout=0.05)
                        except queue.Empty:
                            pass
                    with self._lock:
                        self._in_use += 1
                        self.max_slots_in_use = max(self.max_slots_in_use, self._in_use)
                    f0, f1 = c * self.chunk, min(self.n, (c + 1) * self.chunk)
                    present = np.fromiter(pool.map(self._read, [(slot, i - f0, i) for i in range(f0, f1)]), bool, f1 - f0)
                    self._ready.put((c, slot, f0, f1, present, None))
        except BaseException as e:       # surfaces in the consumer
            self._ready.put((-1, -1, 0, 0, None, e))

    def _release(self, slot):
        with self._lock:
            self._in_use -= 1
        self._free.put(slot)

    # -- consumer -------------------------------------------------------------------------------------------
    def __iter__(self):
        self._thread.start()
        torch = self._torch
        try:
            if torch is None:
                for _ in range(self.n_chunks):
                    c, slot, f0, f1, present, err = self._ready.get()
                    if err is not None:
                        raise err
                    yield f0, f1, self._np_slots[slot][: f1 - f0].copy(), present
                    self._release(slot)
                return
            main = torch.cuda.current_stream(self.device)
            compute_done = [None, None]            # event after the consumer's work on device buffer b was enqueued
            pending = None                         # (chunk info, slot, copy event) whose H2D is in flight

            def start_copy():
                c, slot, f0, f1, present, err = self._ready.get()
                if err is not None:
                    raise err
                b = c % 2
                if compute_done[b] is not None:
                    self.copy_stream.wait_event(compute_done[b])       # the buffer's previous chunk has been consumed
                with torch.cuda.stream(self.copy_stream):
                    self.dev_bufs[b][: f1 - f0].copy_(self.slots[slot][: f1 - f0], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                return (c, f0, f1, present), slot, ev

            if self.n_chunks:
                pending = start_copy()
            for c in range(self.n_chunks):
                (cc, f0, f1, present), slot, ev = pending
                # chunk c + 1 starts travelling before chunk c is handed out (the consumer's call may block)
                pending = start_copy() if c + 1 < self.n_chunks else None
                main.wait_event(ev)
                self.host_view = self._np_slots[slot][: f1 - f0]      # the same frames on the host, valid inside the loop body
                yield f0, f1, self.dev_bufs[cc % 2][: f1 - f0], present
                done = torch.cuda.Event()
                done.record(main)
                compute_done[cc % 2] = done
                ev.synchronize()                   # the pinned slot has been read: back to the ring
                self._release(slot)
        finally:
            self._stop.set()

    def close(self):
        self._stop.set()


def stream_side(depth_data_io, side, dataset, chunk_frames: int = 64, device=None, slots: int = 3) -> RawDepthStreamer:
    """Streamer over the frames of `dataset` (one side of a project)."""
    n = len(dataset.timestamps)
    H, W = (int(dataset.heights[0]), int(dataset.widths[0])) if n else (0, 0)
    if n and not (np.all(np.asarray(dataset.heights) == H) and np.all(np.asarray(dataset.widths) == W)):
        raise RuntimeError("all depth frames of a side must share one resolution")
    ts = np.asarray(dataset.timestamps)
    return RawDepthStreamer(lambda i: depth_data_io.depth_map_path(side, ts[i]), n, H, W, chunk_frames=chunk_frames,
                            slots=slots, device=device)
