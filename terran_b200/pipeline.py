"""Streaming video pipeline: prefetching frame feeder + concurrent detect/pose.

Shape of the reference's end-to-end use (``examples/video.py:33-42``,
``terran/io/video/reader.py:126-162``): a background reader thread prefetches
the next batch of ``rgb24`` frames through a ``Queue(1)`` while the main thread
runs the perception callables on the current one.  Here the prefetch also
covers the host->device copy (pinned staging buffers, a dedicated copy stream,
double buffering), face detection and pose estimation of one batch run
concurrently on two CUDA streams, and the host-side result unpacking of batch i
overlaps the GPU work of batch i+1.
"""
import os
import queue
import threading

import numpy as np
import torch

from terran_b200.defaults import completion_event, cuda_index, default_device


class FrameFeeder:
    """Iterate over batches of uint8 frames, delivering each as a CUDA tensor
    whose upload overlapped the processing of the previous batch.

    ``source`` yields (N,H,W,3) uint8 numpy arrays or CPU tensors (pinned
    tensors are uploaded without a staging copy).  ``depth`` batches are in
    flight (reference: ``Queue(1)`` prefetch)."""

    #: ONE upload stream per device for every feeder of the process.  The caching allocator keeps a
    #: pool of blocks per stream: a feeder with a private stream would find that pool empty and
    #: pay two or three ``cudaMalloc`` of a whole batch (199 MB for 32 x 1080p) before its first
    #: frame is up — a driver call that takes anything from 1 ms to 0.5 s on a shared host
    #: (profiles/r02_e2e_ranks.txt).  With the shared stream a new feeder re-uses the blocks the
    #: previous one released.
    _copy_streams = {}
    _copy_streams_lock = threading.Lock()

    @classmethod
    def copy_stream(cls, device_index):
        with cls._copy_streams_lock:
            stream = cls._copy_streams.get(device_index)
            if stream is None:
                stream = cls._copy_streams[device_index] = torch.cuda.Stream(device=device_index)
            return stream

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
        if os.environ.get('TRB_FEEDER_PRIVATE_STREAM', '0') == '1':      # (A/B switch of the profile above)
            stream = torch.cuda.Stream(device=self.device_index)
        else:
            stream = self.copy_stream(self.device_index)
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
                    done = completion_event()
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
            cur = torch.cuda.current_stream(self.device_index)
            cur.wait_event(done)
            done.synchronize()
            # The batch was allocated on the feeder's private copy stream: tell the caching
            # allocator that the consumer's stream uses it too, so that dropping the tensor while
            # kernels are still queued there cannot hand the block to the next upload.  (A consumer
            # that moves the work to yet another stream must call ``record_stream`` itself, as
            # ``PerceptionPipeline.submit`` does.)
            dev.record_stream(cur)
            yield dev

    def close(self):
        self.stop.set()


class PerceptionPipeline:
    """Face detection + pose estimation of a stream of frame batches.

    ``submit(frames)`` enqueues both tasks for one batch on two CUDA streams
    (they are independent: the small post-processing kernels of one overlap the
    convolutions of the other) without synchronising the host; ``result()`` of
    the returned handle downloads and unpacks.  ``run(batches)`` keeps one batch
    of lookahead: the GPU works on batch i+1 while the host unpacks batch i —
    the synchronous per-batch loop of the reference's ``examples/video.py``,
    software-pipelined."""

    def __init__(self, detection, estimation, device=default_device, tracker=None, recognition=None):
        """tracker: optional ``terran_b200.tracking.Sort``; ``run`` then feeds it the faces of
        every frame in stream order (tracking is the one consumer that needs frame order) and
        yields the faces with their ``track`` field (reference: ``examples/match.py:31-40``)."""
        self.detection, self.estimation = detection, estimation
        #: optional ``Recognition``: every detected face is aligned and embedded on the device
        #: right after detection (reference pipeline shape: ``examples/match.py:29-33``); the
        #: results then are (faces, features_per_image, poses)
        self.recognition = recognition
        self.tracker = tracker
        self.device_index = cuda_index(device)
        self.streams = [torch.cuda.Stream(device=self.device_index) for _ in range(2)]

    def submit(self, frames):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device_index))
        handles = []
        for fn, stream in zip((self.detection.submit, self.estimation.submit), self.streams):
            with torch.cuda.stream(stream):
                stream.wait_event(ready)
                if isinstance(frames, torch.Tensor) and frames.is_cuda:
                    frames.record_stream(stream)     # allocated on the feeder's copy stream
                handles.append(fn(frames))
        if self.recognition is not None:
            handles.insert(1, self._submit_recognition(frames, handles[0]))
        return _PendingPair(handles)

    def _submit_recognition(self, frames, det_handle):
        """Embed the faces of ``det_handle`` on the detection stream.  Frames that did not go
        through the device path (lists of images) fall back to the public call."""
        rec = self.recognition
        if rec.model is None:
            rec.model = rec.recognition_cls(device=rec.device)
        pending = getattr(det_handle, 'pending', None)
        if pending is None or not hasattr(rec.model, 'embed_detections'):
            return _Lazy(lambda: rec(frames, det_handle.result()))
        from terran_b200.frames import to_device_u8
        with torch.cuda.device(self.device_index), torch.cuda.stream(self.streams[0]):
            dev_frames = to_device_u8(frames, self.device_index)
            features, counts = rec.model.embed_detections(dev_frames, pending, det_handle.scale)
            host = torch.empty(features.shape, dtype=torch.float32).pin_memory()
            host.copy_(features, non_blocking=True)
            done = completion_event()
            done.record()

        def finish():
            done.synchronize()
            splits = np.cumsum(counts)[:-1]
            return [np.empty((0, 512)) if not c else part      # (float64 empties, as upstream)
                    for c, part in zip(counts, np.split(host.numpy(), splits, axis=0))]
        return _Lazy(finish)

    def __call__(self, frames):
        return self.submit(frames).result()

    def run(self, batches):
        """Yield (faces, poses) per batch, one batch of lookahead."""
        previous = None
        for frames in batches:
            current = self.submit(frames)
            if previous is not None:
                yield self._tracked(previous.result())
            previous = current
        if previous is not None:
            yield self._tracked(previous.result())

    def _tracked(self, result):
        if self.tracker is None:
            return result
        faces, rest = result[0], result[1:]
        return ([self.tracker.update(f) for f in faces],) + tuple(rest)

    def close(self):
        pass


class _Lazy:
    def __init__(self, fn):
        self._fn, self._done, self._value = fn, False, None

    def result(self):
        if not self._done:
            self._value, self._done, self._fn = self._fn(), True, None
        return self._value


class _PendingPair:
    def __init__(self, handles):
        self.handles = handles

    def result(self):
        return tuple(h.result() for h in self.handles)
