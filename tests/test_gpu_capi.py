"""GPU: the single-call C entry points (``tr_*_create`` from a checkpoint blob, ``tr_*_forward``
on device buffers) — what a non-Python host binds — give exactly what the Python model classes
give (which the other GPU tests check against the oracle)."""
import ctypes as C

import numpy as np
import pytest
import torch

from terran_b200 import synth
from terran_b200.weights import pack_state_dict

pytestmark = pytest.mark.gpu


def create(nat, name, sd):
    blob = np.frombuffer(pack_state_dict(sd), np.uint8)
    h = C.c_void_p()
    nat.init(0)
    nat.check(getattr(nat.lib(), f'tr_{name}_create')(C.c_void_p(blob.ctypes.data), len(blob), C.byref(h)))
    return h


def ptr(t):
    return C.c_void_p(t.data_ptr())


def test_retinaface_single_call(native):
    nat = native
    from terran_b200.face.detection.retinaface import RetinaFace
    sd = synth.retinaface_state_dict()
    frames = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (3, 208, 370, 3), dtype=np.uint8)).cuda()
    want_count, _, want_det = RetinaFace(device=torch.device('cuda'), state_dict=sd).detect_device(frames)
    h = create(nat, 'retinaface', sd)
    try:
        count = torch.empty(3, dtype=torch.int32, device='cuda')
        det = torch.empty((3, 512, 16), dtype=torch.float32, device='cuda')
        for _ in range(2):          # second call re-uses the handle's scratch
            nat.check(nat.lib().tr_retinaface_forward(h, ptr(frames), 3, 208, 370, 0.5, 0.4, 512, ptr(count),
                                                      ptr(det), nat.current_stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(count, want_count) and int(count.sum()) > 0
        for n in range(3):
            k = int(count[n])
            assert torch.equal(det[n, :k].view(torch.int32), want_det[n, :k].view(torch.int32))
    finally:
        nat.lib().tr_model_destroy(h)


def test_arcface_single_call(native):
    nat = native
    from terran_b200.face.recognition.arcface import ArcFace
    sd = synth.arcface_state_dict(units=(1, 1, 1, 1))
    crops = torch.from_numpy(np.random.default_rng(4).integers(0, 256, (5, 112, 112, 3), dtype=np.uint8)).cuda()
    model = ArcFace(device=torch.device('cuda'), state_dict=sd)
    want = model.embed_device(crops)
    want_raw = model.embed_device(crops, normalise=False)
    h = create(nat, 'arcface', sd)
    try:
        emb = torch.empty((5, 512), dtype=torch.float32, device='cuda')
        nat.check(nat.lib().tr_arcface_forward(h, ptr(crops), 5, 0, 1, ptr(emb), nat.current_stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(emb, want)
        nat.check(nat.lib().tr_arcface_forward(h, ptr(crops), 5, 0, 0, ptr(emb), nat.current_stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(emb, want_raw)
        chw = crops.permute(0, 3, 1, 2).flip(1).contiguous()          # the reference's model input
        nat.check(nat.lib().tr_arcface_forward(h, ptr(chw), 5, 1, 1, ptr(emb), nat.current_stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(emb, want)
    finally:
        nat.lib().tr_model_destroy(h)


def test_openpose_single_call(native):
    nat = native
    from terran_b200.pose.openpose import OpenPose
    from terran_b200.pose.openpose.wrapper import parse_device
    sd = synth.openpose_state_dict(peaks=True)
    from terran_b200.frames import resize_short_side
    big = torch.from_numpy(np.random.default_rng(5).integers(0, 256, (2, 720, 1280, 3), dtype=np.uint8)).cuda()
    frames, _ = resize_short_side(big, 184)
    model = OpenPose(device=torch.device('cuda'), state_dict=sd)
    paf, heat = model.maps(frames)
    scale = 184 / 720
    w_count, w_kps, w_score, w_status, _ = parse_device(paf, heat, scale)
    h = create(nat, 'openpose', sd)
    try:
        count = torch.empty(2, dtype=torch.int32, device='cuda')
        kps = torch.empty((2, nat.TR_HUMAN_CAP, 18, 3), dtype=torch.int32, device='cuda')
        score = torch.empty((2, nat.TR_HUMAN_CAP), dtype=torch.float64, device='cuda')
        status = torch.empty(2, dtype=torch.int32, device='cuda')
        nat.check(nat.lib().tr_openpose_forward(h, ptr(frames), 2, 184, 327, scale, ptr(count), ptr(kps),
                                                ptr(score), ptr(status), nat.current_stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(count, w_count) and int(count.sum()) > 0 and not status.any()
        for n in range(2):
            k = int(count[n])
            assert torch.equal(kps[n, :k], w_kps[n, :k]) and torch.equal(score[n, :k], w_score[n, :k])
    finally:
        nat.lib().tr_model_destroy(h)
