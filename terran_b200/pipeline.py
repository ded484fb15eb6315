"""Streaming video pipeline: prefetching frame feeder + concurrent detect/pose.

Shape of the reference's end-to-end use (``examples/video.py:33-42``,
``terran/io/video/reader.py:126-162``): a background reader thread prefetches
the next batch of ``rgb24`` frames through a ``Queue(1)`` while the main thread
runs the perception callables on the current one.  Here the prefetch also
covers the host->device copy (pinned staging buffers, a dedicated copy stream,
double buffering), and face detection and pose estimation of one batch run
concurrently on two CUDA streams driven by two host threads, so the host-side
result unpacking of one task overlaps the GPU work of the other.
"""
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from terran_b200.defaults import cuda_index, default_device


class FrameFeeder:
    """Iterate over batches of uint8 frames, delivering each as a CUDA tensor
    whose upload overlapped the processing of the previous batch.

    ``source`` yields (N,H,W,3) uint8 numpy arrays or CPU tensors (pinned
    tensors are uploaded without a staging copy).  ``depth`` batches are in
    flight (reference: ``Queue(1)`` prefetch)."""

    def __init__(self, source, device=default_device, depth=2):
        self.source = iter(source)
        self.device_index = cuda_index(device)
        self.depth = depth
        self.queue = queue.Queue(maxsize=depth)
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.started = False

    def _run(self):
        torch.cuda.set_device(self.device_index)
        stream = torch.cuda.Stream(device=self.device_index)
        staging = {}
        slot = 0
        try:
            for batch in self.source:
                if self.stop.is_set():
                    break
                t = batch if isinstance(batch, torch.Tensor) else torch.from_numpy(
                    np.ascontiguousarray(batch))
                if not t.is_pinned():
                    key = (slot % (self.depth + 1), tuple(t.shape))
                    if key not in staging:
                        staging[key] = torch.empty(t.shape, dtype=torch.uint8).pin_memory()
                    staging[key].copy_(t)
                    t = staging[key]
                    slot += 1
                with torch.cuda.stream(stream):
                    dev = t.to(f'cuda:{self.device_index}', non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(stream)
                self.queue.put((dev, done))
        finally:
            self.queue.put(None)

    def __iter__(self):
        if not self.started:
            self.thread.start()
            self.started = True
        while True:
            item = self.queue.get()
            if item is None:
                return
            dev, done = item
            # consumers on any stream must see the finished copy
            torch.cuda.current_stream(self.device_index).wait_event(done)
            done.synchronize()
            yield dev

    def close(self):
        self.stop.set()


class PerceptionPipeline:
    """``faces, poses = pipeline(frames)``: the two public callables on the same
    batch, concurrently (one host thread + one CUDA stream each)."""

    def __init__(self, detection, estimation, device=default_device):
        self.detection, self.estimation = detection, estimation
        self.device_index = cuda_index(device)
        self.pool = ThreadPoolExecutor(max_workers=2)
        self.streams = [torch.cuda.Stream(device=self.device_index) for _ in range(2)]

    def _call(self, fn, stream, frames, ready):
        torch.cuda.set_device(self.device_index)
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            if isinstance(frames, torch.Tensor) and frames.is_cuda:
                frames.record_stream(stream)    # allocated on the feeder's copy stream
            out = fn(frames)
            stream.synchronize()
        return out

    def __call__(self, frames):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device_index))
        a = self.pool.submit(self._call, self.detection, self.streams[0], frames, ready)
        b = self.pool.submit(self._call, self.estimation, self.streams[1], frames, ready)
        return a.result(), b.result()

    def close(self):
        self.pool.shutdown()
