"""OpenPose (2017 body model) — the plugin ``Estimation`` calls.

Same protocol as the reference's ``OpenPose`` class
(``terran/pose/openpose/wrapper.py:166-485``): ``cls(device=..., short_side=184)``
and ``call(images (N,H,W,3) uint8) -> list[list[{'keypoints': int32 (18,3),
'score': float64}]]``.  Resize, VGG/PAF forward, the x8 bicubic up-sampling,
peak extraction, PAF line integrals, greedy limb matching and human assembly
all run in the native library; one D2H copy of the assembled humans per batch
replaces the reference's per-candidate ``.cpu()`` calls.
"""
import ctypes as C

import numpy as np
import torch

from terran_b200 import _native as nat
from terran_b200.checkpoint import get_checkpoint_path
from terran_b200.defaults import completion_event, cuda_index, default_device
from terran_b200.frames import resize_short_side, to_device_u8
from terran_b200.weights import Net, openpose_program

CLASS_PATH = 'terran_b200.pose.openpose.OpenPose'


def parse_device(paf, heat, scale, workspace=None):
    """Stage-level entry: network maps (CUDA fp32 NCHW, paf (N,38,h,w), heat
    (N,19,h,w)) -> (count, keypoints, score, status) CUDA tensors."""
    N, _, h, w = paf.shape
    dev = paf.device
    nat.init(dev.index or 0)
    paf, heat = paf.contiguous().float(), heat.contiguous().float()
    need = nat.lib().tr_pose_workspace_bytes(N)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    count = torch.empty(N, dtype=torch.int32, device=dev)
    kps = torch.empty((N, nat.TR_HUMAN_CAP, 18, 3), dtype=torch.int32, device=dev)
    score = torch.empty((N, nat.TR_HUMAN_CAP), dtype=torch.float64, device=dev)
    status = torch.empty(N, dtype=torch.int32, device=dev)
    nat.check(nat.lib().tr_openpose_parse(
        C.c_void_p(paf.data_ptr()), C.c_void_p(heat.data_ptr()), N, h, w, float(scale),
        C.c_void_p(workspace.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(kps.data_ptr()),
        C.c_void_p(score.data_ptr()), C.c_void_p(status.data_ptr()), nat.current_stream_ptr()))
    return count, kps, score, status, workspace


def unpack_poses(count, kps, score, status):
    """Device or (pinned) host tensors -> the reference's list of lists of dicts."""
    count, status = count.cpu().numpy().copy(), status.cpu().numpy().copy()
    if status.any():
        # The reference has no capacity limits and returns something for every frame: one
        # crowded / noisy frame must not fail the batch.  The kernels clamp at their capacities
        # (never write out of bounds), so the frame's result is TRUNCATED, and we say so.
        import warnings
        bad = np.flatnonzero(status)
        warnings.warn(
            f'pose parse capacity exceeded on frame(s) {bad.tolist()} (status '
            f'{[int(status[b]) for b in bad]}: 1 = more than {nat.TR_PEAK_CAP} peaks of one part, '
            f'2 = more than {nat.TR_CAND_CAP} limb candidates, 4 = more than {nat.TR_HUMAN_CAP} '
            f'humans): results of these frames are truncated', RuntimeWarning, stacklevel=2)
    top = int(count.max()) if len(count) else 0
    kps = kps[:, :max(top, 1)].cpu().numpy().copy()
    score = score[:, :max(top, 1)].cpu().numpy().copy()
    # (kps / score are fresh per-call copies: the per-person arrays are views into them)
    return [
        [{'keypoints': kp, 'score': sc} for kp, sc in zip(kps[n, :k], score[n, :k])]
        for n, k in enumerate(count)
    ]


class OpenPose:

    def __init__(self, device=default_device, short_side=184, state_dict=None):
        self.device = device
        self.device_index = cuda_index(device)
        self.short_side = short_side
        self.downsampling_ratio = 8
        self.keypoint_threshold = 0.1
        self.thresh_2 = 0.05
        self.human_threshold = 0.4
        if state_dict is None:
            state_dict = torch.load(get_checkpoint_path(CLASS_PATH), map_location='cpu')
        program, self.roles = openpose_program(state_dict)
        with torch.cuda.device(self.device_index):
            self.net = Net(program, self.device_index)
        self._ws = None
        self._host = {}
        self._parse_stream = None

    # -- device stages --------------------------------------------------------
    def maps(self, resized):
        """resized: CUDA uint8 (N,h,w,3) RGB at network resolution.  Returns the
        reference module's outputs (paf (N,38,h/8,w/8), heat (N,19,...)) fp32."""
        N, H, W, _ = resized.shape
        self.net.run(resized, N, H, W, (H * W * 3, W * 3, 3, 1))
        r = self.roles
        return (self.net.export_nchw(r['maps'], r['paf_coff'], 38),
                self.net.export_nchw(r['maps'], r['heat_coff'], 19))

    def estimate_device(self, frames):
        with nat.nvtx_range('pose:resize'):
            resized, scale = resize_short_side(frames, self.short_side)
        with nat.nvtx_range('pose:net'):
            paf, heat = self.maps(resized)
        with nat.nvtx_range('pose:parse'):
            count, kps, score, status, self._ws = parse_device(paf, heat, scale, self._ws)
        return count, kps, score, status

    def estimate_async(self, frames):
        """Enqueue resize + forward on the current stream and parse + the D2H of the results
        on the model's own side stream, without host synchronisation; returns a ``PendingPoses``.
        The parse kernels (peaks, limbs, greedy matching, assembly) are latency-bound and occupy
        few SMs: on their own stream they overlap the convolutions of whatever the caller
        enqueues next (the next batch) instead of idling the GPU behind the conv stack."""
        with nat.nvtx_range('pose:resize'):
            resized, scale = resize_short_side(frames, self.short_side)
        with nat.nvtx_range('pose:net'):
            paf, heat = self.maps(resized)
        cur = torch.cuda.current_stream()
        if self._parse_stream is None:
            self._parse_stream = torch.cuda.Stream(device=self.device_index)
        ready = torch.cuda.Event()
        ready.record(cur)
        N = frames.shape[0]
        with torch.cuda.stream(self._parse_stream), nat.nvtx_range('pose:parse'):
            self._parse_stream.wait_event(ready)
            paf.record_stream(self._parse_stream)
            heat.record_stream(self._parse_stream)
            count, kps, score, status, self._ws = parse_device(paf, heat, scale, self._ws)
            return self._download(N, count, kps, score, status)

    def _download(self, N, count, kps, score, status):
        ring = self._host.setdefault(N, {'i': 0, 'slots': []})
        if len(ring['slots']) < 3:
            ring['slots'].append({
                'count': torch.empty(N, dtype=torch.int32).pin_memory(),
                'status': torch.empty(N, dtype=torch.int32).pin_memory(),
                'kps': torch.empty((N, nat.TR_HUMAN_CAP, 18, 3), dtype=torch.int32).pin_memory(),
                'score': torch.empty((N, nat.TR_HUMAN_CAP), dtype=torch.float64).pin_memory()})
            ring['i'] = len(ring['slots']) - 1
            slot = ring['slots'][-1]
        else:
            ring['i'] = (ring['i'] + 1) % 3
            slot = ring['slots'][ring['i']]
        slot['count'].copy_(count, non_blocking=True)
        slot['status'].copy_(status, non_blocking=True)
        slot['kps'].copy_(kps, non_blocking=True)
        slot['score'].copy_(score, non_blocking=True)
        done = completion_event()
        done.record()
        return PendingPoses(slot, done)

    def call(self, images):
        """Pose estimation on a (N,H,W,3) uint8 RGB batch (numpy or CUDA tensor)."""
        with torch.cuda.device(self.device_index):
            frames = to_device_u8(images, self.device_index)
            return self.estimate_async(frames).result()


class PendingPoses:
    """Results of ``OpenPose.estimate_async`` still in flight."""

    def __init__(self, slot, done):
        self.slot, self.done = slot, done

    def result(self):
        self.done.synchronize()
        s = self.slot
        return unpack_poses(s['count'], s['kps'], s['score'], s['status'])
