"""RetinaFace (mnet) model class — the plugin the ``Detection`` wrapper calls.

Same protocol as the reference's ``RetinaFace`` class
(``terran/face/detection/retinaface/wrapper.py:92-238``): ``cls(device=...)``
and ``call(images, threshold=0.5) -> list[list[dict]]`` with float32 ``bbox``
(4,), ``landmarks`` (5,2) and ``score``.  The network forward, anchor decode,
threshold, sort and NMS all run in the native library; one D2H copy per batch
replaces the reference's three per image.
"""
import ctypes as C

import numpy as np
import torch

from terran_b200 import _native as nat
from terran_b200.checkpoint import get_checkpoint_path
from terran_b200.defaults import completion_event, cuda_index, default_device
from terran_b200.frames import to_device_u8
from terran_b200.weights import Net, retinaface_program

CLASS_PATH = 'terran_b200.face.detection.retinaface.RetinaFace'


def load_state_dict():
    return torch.load(get_checkpoint_path(CLASS_PATH), map_location='cpu')


class RetinaFace:

    def __init__(self, device=default_device, nms_threshold=0.4, state_dict=None):
        self.device = device
        self.device_index = cuda_index(device)
        self.nms_threshold = nms_threshold
        self.feature_strides = [32, 16, 8]
        if state_dict is None:
            state_dict = load_state_dict()
        program, self.roles = retinaface_program(state_dict)
        with torch.cuda.device(self.device_index):
            self.net = Net(program, self.device_index)
        self._ws = None
        self._host = {}
        self.last_candidates = None

    # -- stages, exposed for the parity tests --------------------------------
    def forward(self, frames):
        """frames: CUDA uint8 (N,H,W,3) RGB.  Runs the conv stack."""
        N, H, W, _ = frames.shape
        # model channel order is BGR: start at channel 2, walk backwards
        self.net.run(frames, N, H, W, (H * W * 3, W * 3, 3, -1), ptr_offset=2)

    def heads(self, frames):
        """The reference module's 9 outputs (s32,s16,s8 x prob/bbox/landmark) as
        NCHW fp32 CUDA tensors — for parity against ``RetinaFace.forward``."""
        self.forward(frames)
        out = []
        for buf in self.roles['heads']:
            out.append(self.net.export_nchw_f32(buf, 0, 4, softmax_pairs=True))
            out.append(self.net.export_nchw_f32(buf, 4, 8))
            out.append(self.net.export_nchw_f32(buf, 12, 20))
        return out

    def _workspace(self, N, H, W):
        need = nat.lib().tr_detect_workspace_bytes(N, H, W)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=f'cuda:{self.device_index}')
        return self._ws

    def detect_device(self, frames, threshold=0.5, max_det=512):
        """Forward + post-processing, results left on the device.
        Returns (count (N,), candidates (N,), det (N,max_det,16))."""
        N, H, W, _ = frames.shape
        self.forward(frames)
        dev = frames.device
        ws = self._workspace(N, H, W)
        while True:
            count = torch.empty(N, dtype=torch.int32, device=dev)
            cand = torch.empty(N, dtype=torch.int32, device=dev)
            det = torch.empty((N, max_det, 16), dtype=torch.float32, device=dev)
            heads = (C.c_int * 3)(*self.roles['heads'])
            nat.check(nat.lib().tr_retinaface_detect(
                self.net.handle, heads, float(threshold), float(self.nms_threshold), max_det,
                C.c_void_p(ws.data_ptr()), C.c_void_p(count.data_ptr()),
                C.c_void_p(cand.data_ptr()), C.c_void_p(det.data_ptr()),
                nat.current_stream_ptr()))
            top = int(count.max().item()) if N else 0
            if top <= max_det:
                return count, cand, det
            max_det = top            # rare: more survivors than rows; redo the select

    def detect_async(self, frames, threshold=0.5, max_det=512):
        """Enqueue forward + post-processing + the D2H of the results on the current
        stream without any host synchronisation; returns a ``PendingDetections``."""
        N, H, W, _ = frames.shape
        dev = frames.device
        with nat.nvtx_range('detect:net'):
            self.forward(frames)
        ws = self._workspace(N, H, W)
        count = torch.empty(N, dtype=torch.int32, device=dev)
        cand = torch.empty(N, dtype=torch.int32, device=dev)
        det = torch.empty((N, max_det, 16), dtype=torch.float32, device=dev)
        heads = (C.c_int * 3)(*self.roles['heads'])
        with nat.nvtx_range('detect:decode+nms'):
            nat.check(nat.lib().tr_retinaface_detect(
                self.net.handle, heads, float(threshold), float(self.nms_threshold), max_det,
                C.c_void_p(ws.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(cand.data_ptr()),
                C.c_void_p(det.data_ptr()), nat.current_stream_ptr()))
        slot = self._host_slot(N, max_det)
        slot['count'].copy_(count, non_blocking=True)
        slot['det'].copy_(det, non_blocking=True)
        done = completion_event()
        done.record()
        self.last_candidates = cand
        return PendingDetections(self, frames, threshold, max_det, slot, done, count, det)

    def _host_slot(self, N, max_det):
        """Pinned result buffers, three sets in rotation (a result may still be
        unread while the next two batches are in flight)."""
        key = (N, max_det)
        ring = self._host.setdefault(key, {'i': 0, 'slots': []})
        if len(ring['slots']) < 3:
            ring['slots'].append({
                'count': torch.empty(N, dtype=torch.int32).pin_memory(),
                'det': torch.empty((N, max_det, 16), dtype=torch.float32).pin_memory()})
            ring['i'] = len(ring['slots']) - 1
            return ring['slots'][-1]
        ring['i'] = (ring['i'] + 1) % 3
        return ring['slots'][ring['i']]

    def call_arrays(self, images, threshold=0.5):
        """``call`` without the per-face dict building: (counts (N,), rows
        (N,R,16) float32 on the host) — see ``unpack_detections``."""
        with torch.cuda.device(self.device_index):
            frames = to_device_u8(images, self.device_index)
            return self.detect_async(frames, threshold).arrays()

    def call(self, images, threshold=0.5):
        """Run the detection.  ``images`` is a (N,H,W,3) uint8 RGB array
        (numpy, or a CUDA tensor to skip the upload)."""
        return unpack_detections(*self.call_arrays(images, threshold))


class PendingDetections:
    """Results of ``RetinaFace.detect_async`` still in flight."""

    def __init__(self, model, frames, threshold, max_det, slot, done, count_dev=None, det_dev=None):
        self.model, self.frames, self.threshold = model, frames, threshold
        self.max_det, self.slot, self.done = max_det, slot, done
        #: device copies of (count (N,), rows (N,max_det,16)) for consumers that stay on the
        #: GPU (``ArcFace.embed_detections``: detect -> align -> embed without a host round trip)
        self.count_dev, self.det_dev = count_dev, det_dev
        self.stream = torch.cuda.current_stream()

    def arrays(self):
        """(counts (N,), rows (N,R,16)) on the host; blocks until the copy landed."""
        self.done.synchronize()
        counts = self.slot['count'].numpy().copy()
        top = int(counts.max()) if len(counts) else 0
        if top > self.max_det:
            # rare: more survivors than rows were copied — redo synchronously with room
            with torch.cuda.device(self.model.device_index), torch.cuda.stream(self.stream):
                # (on the stream the batch was submitted on: the net's activation buffers are
                # shared with whatever that stream runs next)
                count, _, det = self.model.detect_device(self.frames, self.threshold, max_det=top)
                self.count_dev, self.det_dev, self.max_det = count, det, top
                return count.cpu().numpy(), det.cpu().numpy()
        return counts, self.slot['det'][:, :max(top, 1)].numpy().copy()

    def result(self, scale=None):
        return unpack_detections(*self.arrays(), scale=scale)


def unpack_detections(counts, rows, scale=None):
    """(N,) counts + (N,R,16) rows -> the reference's list of lists of dicts.
    With ``scale`` the coordinates are mapped back to input pixels exactly like
    ``Detection.resize_out`` (detection/__init__.py:59-84): float32 value /
    python-float scale, round half to even, int32 — done once on the whole
    batch instead of per face."""
    scores = np.ascontiguousarray(rows[..., 0])
    coords = rows[..., 1:15]
    if scale is not None:
        with np.errstate(invalid='ignore', over='ignore'):       # rows past `count` are uninitialised
            coords = np.around(coords / scale).astype(np.int32)
    else:
        coords = np.ascontiguousarray(coords)
    boxes = coords[..., :4]
    lmks = coords[..., 4:].reshape(coords.shape[0], coords.shape[1], 5, 2)
    # (iterating an array yields its rows as views at C speed: ~2x faster than indexing [n, i])
    return [
        [{'bbox': b, 'landmarks': l, 'score': s}
         for b, l, s in zip(boxes[n, :k], lmks[n, :k], scores[n, :k])]
        for n, k in enumerate(counts)
    ]


def decode_nms(heads, H, W, threshold=0.5, nms_threshold=0.4, max_det=None):
    """Stage-level entry: the reference's 9 head tensors (torch CUDA fp32 NCHW,
    class scores already soft-maxed) -> (counts, candidates, rows, indices)."""
    N = heads[0].shape[0]
    dev = heads[0].device
    nat.init(dev.index or 0)
    heads = [h.contiguous().float() for h in heads]
    A = sum(int(h.shape[2] * h.shape[3] * 2) for h in heads[::3])
    max_det = A if max_det is None else max_det
    ws = torch.empty(nat.lib().tr_detect_workspace_bytes(N, H, W), dtype=torch.uint8, device=dev)
    count = torch.empty(N, dtype=torch.int32, device=dev)
    cand = torch.empty(N, dtype=torch.int32, device=dev)
    det = torch.empty((N, max_det, 16), dtype=torch.float32, device=dev)
    ptrs = (C.c_void_p * 9)(*[h.data_ptr() for h in heads])
    nat.check(nat.lib().tr_retinaface_decode_nms(
        ptrs, N, H, W, float(threshold), float(nms_threshold), max_det,
        C.c_void_p(ws.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(cand.data_ptr()),
        C.c_void_p(det.data_ptr()), nat.current_stream_ptr()))
    counts = count.cpu().numpy()
    rows = det.cpu().numpy()
    idx = rows[..., 15].copy().view(np.int32)
    return counts, cand.cpu().numpy(), rows, idx
