"""CPU: the oracle restatements against the golden fixtures that
``oracle/make_golden.py`` produced by running the UNMODIFIED reference
(``/root/reference``) on the same seeded inputs and synthetic checkpoints."""
import numpy as np
import pytest
import torch

from oracle import detect, nets, pose
from terran_b200 import synth


@pytest.fixture(scope='module')
def retina_sd():
    return synth.retinaface_state_dict()


def test_retinaface_forward_matches_reference(golden, retina_sd):
    g = golden('retinaface_forward.npz')
    x = torch.from_numpy(g['x'].astype(np.float32))
    heads = nets.retinaface_forward(retina_sd, x)
    assert len(heads) == 9
    for i, h in enumerate(heads):
        np.testing.assert_allclose(h.numpy(), g[f'head{i}'], rtol=0, atol=2e-5)


def test_anchor_reference_closed_form():
    # generate_anchors(base 16, ratio 1, scales ...) of anchors.py:75-134
    assert detect.anchor_refs_from_settings(16, (32, 16)) == detect.ANCHOR_REFS[32]
    assert detect.anchor_refs_from_settings(16, (8, 4)) == detect.ANCHOR_REFS[16]
    assert detect.anchor_refs_from_settings(16, (2, 1)) == detect.ANCHOR_REFS[8]
    a = detect.anchors_for(16, 2, 3)
    assert a.shape == (12, 4)
    np.testing.assert_array_equal(a[0], [-56, -56, 71, 71])
    np.testing.assert_array_equal(a[1], [-24, -24, 39, 39])
    np.testing.assert_array_equal(a[2], [-40, -56, 87, 71])      # w = 1
    np.testing.assert_array_equal(a[6], [-56, -40, 71, 87])      # h = 1


def test_retinaface_call_matches_reference(golden):
    """decode + threshold + sort + NMS on the reference's own head tensors:
    identical survivor sets and order, scores bit-equal, boxes within the 1-ulp
    exp() difference (torch's SLEEF exp vs the correctly rounded oracle exp)."""
    g = golden('retinaface_call.npz')
    heads = [g[f'head{i}'] for i in range(9)]
    H, W = g['images'].shape[1:3]
    out = detect.model_call(heads, H, W)
    assert len(out) == g['images'].shape[0]
    for i, faces in enumerate(out):
        ref_scores = g[f'score{i}']
        assert len(faces) == len(ref_scores) > 0
        np.testing.assert_array_equal(np.array([f['score'] for f in faces]), ref_scores)
        np.testing.assert_allclose(np.stack([f['bbox'] for f in faces]), g[f'bbox{i}'],
                                   rtol=0, atol=2e-4)
        np.testing.assert_allclose(np.stack([f['landmarks'] for f in faces]), g[f'landmarks{i}'],
                                   rtol=0, atol=2e-4)


def test_retinaface_full_pipeline_from_pixels(golden, retina_sd):
    """uint8 frames -> oracle network -> oracle post-processing equals the
    reference's RetinaFace.call output."""
    g = golden('retinaface_call.npz')
    images = g['images']
    x = torch.from_numpy(images.astype(np.float32)).permute(0, 3, 1, 2).flip(1)
    heads = [h.numpy() for h in nets.retinaface_forward(retina_sd, x)]
    out = detect.model_call(heads, *images.shape[1:3])
    for i, faces in enumerate(out):
        np.testing.assert_array_equal(np.array([f['score'] for f in faces]), g[f'score{i}'])


def test_detection_640_end_to_end(golden, retina_sd):
    """BASELINE config 1: Detection()(one 640x640 image) through cv2 resize,
    network, post-processing, rescale + round (int32 outputs exact)."""
    import cv2
    g = golden('retinaface_detection_640.npz')
    img = np.random.default_rng(0).integers(0, 256, (640, 640, 3), dtype=np.uint8)
    scale = 416 / 640
    small = cv2.resize(img, (int(640 * scale), int(640 * scale)), interpolation=cv2.INTER_LINEAR)
    x = torch.from_numpy(small[None].astype(np.float32)).permute(0, 3, 1, 2).flip(1)
    heads = [h.numpy() for h in nets.retinaface_forward(retina_sd, x)]
    faces = detect.resize_out(detect.model_call(heads, 416, 416), scale)[0]
    assert len(faces) == len(g['score'])
    np.testing.assert_array_equal(np.stack([f['bbox'] for f in faces]), g['bbox'])
    np.testing.assert_array_equal(np.stack([f['landmarks'] for f in faces]), g['landmarks'])
    np.testing.assert_array_equal(np.array([f['score'] for f in faces]), g['score'])
    assert faces[0]['bbox'].dtype == np.int32


def test_resize_oracle_matches_cv2():
    """The integer restatement of cv2.resize(INTER_LINEAR) on uint8 (the
    reference's host resize) — down- and up-scaling, odd sizes."""
    import cv2
    from oracle import resize
    rng = np.random.default_rng(0)
    for (H, W), short in (((1080, 1920), 416), ((720, 1280), 184), ((640, 640), 416),
                          ((333, 517), 416), ((97, 61), 184), ((50, 70), 416), ((1200, 777), 300)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        got, scale = resize.resize_short_side(img, short)
        want = cv2.resize(img, (int(W * scale), int(H * scale)), interpolation=cv2.INTER_LINEAR)
        np.testing.assert_array_equal(got, want, err_msg=f'{H}x{W}->{short}')


def test_nms_matches_torchvision():
    """The oracle NMS against torchvision.ops.nms (the reference's dependency),
    including the IoU == threshold edge (suppressed only when strictly greater
    as a double)."""
    from torchvision.ops import nms as tv_nms
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = 300
        xy = rng.uniform(0, 200, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 80, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.permutation(n).astype(np.float32) / n
        order = np.argsort(-scores, kind='stable')
        keep = order[detect.nms(boxes[order], 0.4)]
        ref = tv_nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.4).numpy()
        np.testing.assert_array_equal(keep, ref)


def test_nms_edge_cases():
    assert len(detect.nms(np.zeros((0, 4), np.float32), 0.4)) == 0
    one = np.array([[0, 0, 10, 10]], np.float32)
    np.testing.assert_array_equal(detect.nms(one, 0.4), [0])
    # identical boxes: IoU 1 -> second suppressed; degenerate zero-area boxes: IoU nan -> kept
    same = np.array([[0, 0, 10, 10], [0, 0, 10, 10]], np.float32)
    np.testing.assert_array_equal(detect.nms(same, 0.4), [0])
    degenerate = np.array([[5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    np.testing.assert_array_equal(detect.nms(degenerate, 0.4), [0, 1])


def test_arcface_matches_reference(golden):
    g = golden('arcface_embed.npz')
    sd = synth.arcface_state_dict()
    crops = g['crops']
    x = torch.from_numpy(crops.transpose(0, 3, 1, 2)[:, ::-1].astype(np.float32).copy())
    raw = nets.arcface_forward(sd, x).numpy()
    np.testing.assert_allclose(raw, g['raw'], rtol=0, atol=1e-4)
    norm = np.sqrt((raw ** 2).sum(1, keepdims=True))
    norm[norm == 0] = 1
    np.testing.assert_allclose(raw / norm, g['normalised'], rtol=0, atol=1e-6)


def test_openpose_forward_matches_reference(golden):
    g = golden('openpose_forward.npz')
    sd = synth.openpose_state_dict()
    paf, heat = nets.openpose_forward(sd, torch.from_numpy(g['x']))
    np.testing.assert_allclose(paf.numpy(), g['paf'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(heat.numpy(), g['heat'], rtol=0, atol=1e-5)
    # quirk: the final heat-map layer keeps its ReLU (openpose/model.py:38)
    assert heat.min() >= 0 and paf.min() < 0


def test_bicubic_matches_torch():
    t = torch.from_numpy(np.random.default_rng(3).random((1, 4, 9, 13)).astype(np.float32))
    ref = torch.nn.functional.interpolate(t, scale_factor=8, mode='bicubic',
                                          align_corners=False)[0].numpy()
    np.testing.assert_allclose(pose.bicubic_up8(t[0].numpy()), ref, rtol=0, atol=1e-6)
    tab = pose.bicubic_table()
    np.testing.assert_allclose(tab.sum(1), 1.0, atol=1e-6)


def test_openpose_parse_matches_reference(golden):
    """24 synthetic scenes through the reference's OpenPose.call (stub network):
    identical keypoint integers and human counts, scores to 1e-5."""
    g = golden('openpose_parse.npz')
    scale = 184 / 720
    total = 0
    for k, seed in enumerate(g['seeds']):
        paf, heat = pose.synthetic_scene(int(seed))
        np.testing.assert_allclose([paf.astype(np.float64).sum(), heat.astype(np.float64).sum()],
                                   g[f'mapsum{k}'], rtol=1e-12)
        if k < 4:
            np.testing.assert_array_equal(paf, g[f'paf{k}'].astype(np.float32))
            np.testing.assert_array_equal(heat, g[f'heat{k}'].astype(np.float32))
        humans = pose.parse_frame(paf, heat, scale)
        assert len(humans) == len(g[f'score{k}'])
        if humans:
            np.testing.assert_array_equal(np.stack([h['keypoints'] for h in humans]), g[f'kp{k}'])
            np.testing.assert_allclose([h['score'] for h in humans], g[f'score{k}'], atol=1e-5)
        total += len(humans)
    assert total > 50


def test_openpose_parse_edge_cases():
    z_paf, z_heat = np.zeros((38, 6, 7), np.float32), np.zeros((19, 6, 7), np.float32)
    assert pose.parse_frame(z_paf, z_heat, 1.0) == []          # no peaks at all
    # one isolated joint: a peak but no limb -> no human
    h = z_heat.copy()
    h[0, 3, 3] = 1.0
    assert pose.parse_frame(z_paf, h, 1.0) == []
    # zero-length pair (same location for src/dst part) -> NaN score, rejected
    assert pose.segment_points(5, 5) == [5] * 10


# ---- fixtures at the BASELINE batch sizes (oracle/make_golden.py::golden_baseline)

def test_baseline_c3_oracle_matches_reference_256_crops(golden):
    """First rows of the reference's 256-crop embedding batch (the rest is the same
    function of independent crops)."""
    g = golden('arcface_embed_b256.npz')
    crops = np.random.default_rng(int(g['crops_seed'])).integers(0, 256, (256, 112, 112, 3),
                                                                 dtype=np.uint8)
    sel = [0, 100, 255]
    x = torch.from_numpy(crops[sel].transpose(0, 3, 1, 2)[:, ::-1].astype(np.float32).copy())
    raw = nets.arcface_forward(synth.arcface_state_dict(), x).numpy()
    want = g['normalised']
    assert want.shape == (256, 512)
    np.testing.assert_allclose(raw / np.linalg.norm(raw, axis=1, keepdims=True), want[sel],
                               rtol=0, atol=1e-6)


def test_baseline_c4_oracle_estimation_with_humans_matches_reference(golden):
    """The reference's own ``Estimation`` on 720p noise frames with the peak-calibrated
    checkpoint finds 6 and 9 humans; the oracle (cv2 resize -> net -> parse) reproduces
    them exactly."""
    import cv2
    g = golden('openpose_estimation_720p.npz')
    frames = np.random.default_rng(int(g['frames_seed'])).integers(0, 256, (2, 720, 1280, 3),
                                                                   dtype=np.uint8)
    sd = synth.openpose_state_dict(peaks=True)
    s = 184 / 720
    small = np.stack([cv2.resize(f, (int(1280 * s), int(720 * s)), interpolation=cv2.INTER_LINEAR)
                      for f in frames[:1]])
    x = torch.from_numpy(small.transpose(0, 3, 1, 2).astype(np.float32) / 255.0 - 0.5)
    paf, heat = nets.openpose_forward(sd, x)
    out = pose.parse(paf.numpy(), heat.numpy(), s)[0]
    assert len(out) == len(g['kp0']) >= 4
    for o, k, sc in zip(out, g['kp0'], g['score0']):
        np.testing.assert_array_equal(o['keypoints'], k)
        assert abs(o['score'] - sc) < 1e-5


def test_baseline_c2_oracle_detection_1080p_matches_reference(golden, retina_sd):
    import cv2
    g = golden('retinaface_detection_1080p.npz')
    frames = np.random.default_rng(int(g['frames_seed'])).integers(0, 256, (4, 1080, 1920, 3),
                                                                   dtype=np.uint8)[:2]
    s = 416 / 1080
    small = np.stack([cv2.resize(f, (int(1920 * s), int(1080 * s)), interpolation=cv2.INTER_LINEAR)
                      for f in frames])
    x = torch.from_numpy(small.astype(np.float32)).permute(0, 3, 1, 2).flip(1)
    heads = [h.numpy() for h in nets.retinaface_forward(retina_sd, x)]
    out = detect.resize_out(detect.model_call(heads, *small.shape[1:3]), s)
    for n, faces in enumerate(out):
        assert len(faces) == len(g[f'score{n}']) > 10
        np.testing.assert_array_equal(np.stack([f['bbox'] for f in faces]), g[f'bbox{n}'])
        np.testing.assert_array_equal(np.stack([f['landmarks'] for f in faces]), g[f'landmarks{n}'])
        np.testing.assert_array_equal(np.array([f['score'] for f in faces]), g[f'score{n}'])


def test_pil_resize_restatement_matches_pil():
    """oracle/letterbox.py restates Pillow's 8-bit antialiased bicubic resampler: pinned bit-exact
    against PIL itself (the routine the reference calls in preprocess_face_no_landmarks,
    arcface/wrapper.py:75-99) over down- and up-scaling, extreme aspect ratios and tiny images."""
    from PIL import Image
    from oracle import letterbox as lb
    rng = np.random.default_rng(5)
    shapes = [(int(h), int(w)) for h, w in rng.integers(3, 400, (25, 2))]
    shapes += [(112, 112), (1, 1), (2, 300), (300, 2), (720, 1280), (1001, 37)]
    for h, w in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        s = 112 / max(w, h)
        ow, oh = int(w * s), int(h * s)
        if ow < 1 or oh < 1:
            continue
        want = np.asarray(Image.fromarray(img).resize((ow, oh)))
        np.testing.assert_array_equal(lb.pil_resize_bicubic(img, ow, oh), want, err_msg=f'{h}x{w}')
        np.testing.assert_array_equal(lb.letterbox(img), lb.preprocess_face_no_landmarks(img))


def test_letterbox_oracle_matches_reference(golden):
    """oracle/letterbox.py against the outputs of the UNMODIFIED reference's
    preprocess_face_no_landmarks (tests/golden/letterbox.npz, oracle/make_golden.py letterbox)."""
    from oracle import letterbox as lb
    g = golden('letterbox.npz')
    for i in range(int(g['n'])):
        np.testing.assert_array_equal(lb.letterbox(g[f'image_{i}']), g[f'crop_{i}'])
        np.testing.assert_array_equal(lb.preprocess_face_no_landmarks(g[f'image_{i}']), g[f'crop_{i}'])
