"""GPU: the post-processing kernels (RetinaFace decode/sort/NMS, OpenPose
parse) against the numpy oracle on identical fp32 inputs — integer outputs and
survivor sets bit-exact, through the C ABI stage entry points."""
import numpy as np
import pytest
import torch

from oracle import detect, pose

pytestmark = pytest.mark.gpu


def random_heads(rng, N, H, W, frac=0.02, spread=6.0):
    """Reference-layout head tensors with ~frac of anchors above 0.5."""
    heads = []
    for stride in (32, 16, 8):
        fh, fw = -(-H // stride), -(-W // stride)
        logit = rng.normal(-spread * 0.35 - 2.0, spread, (N, 2, fh, fw))
        shift = np.quantile(logit, 1 - frac)
        p_fg = 1 / (1 + np.exp(-(logit - shift)))
        prob = np.concatenate([1 - p_fg, p_fg], 1).astype(np.float32)
        bbox = rng.normal(0, 0.4, (N, 8, fh, fw)).astype(np.float32)
        lmk = rng.normal(0, 0.3, (N, 20, fh, fw)).astype(np.float32)
        heads += [prob, bbox, lmk]
    return heads


def run_native(heads, H, W, thr=0.5, nms_thr=0.4):
    from terran_b200.face.detection.retinaface.wrapper import decode_nms
    dev = [torch.from_numpy(h).cuda() for h in heads]
    return decode_nms(dev, H, W, thr, nms_thr)


def check_against_oracle(heads, H, W, thr=0.5, nms_thr=0.4):
    counts, cands, rows, idx = run_native(heads, H, W, thr, nms_thr)
    s, b, l = detect.decode(heads, H, W)
    ref = detect.select(s, b, l, thr, nms_thr)
    for n, r in enumerate(ref):
        assert cands[n] == r['num_candidates'], (n, cands[n], r['num_candidates'])
        assert counts[n] == len(r['index']), (n, counts[n], len(r['index']))
        k = counts[n]
        np.testing.assert_array_equal(idx[n, :k], r['index'])              # survivor set + order
        np.testing.assert_array_equal(rows[n, :k, 0], r['score'])          # scores bit-exact
        np.testing.assert_array_equal(rows[n, :k, 1:5], r['bbox'])         # boxes bit-exact
        np.testing.assert_array_equal(rows[n, :k, 5:15].reshape(k, 5, 2), r['landmarks'])
    return counts, cands


def test_decode_nms_golden_heads(native, golden):
    """The reference's own head tensors (fixture) -> same survivors as the
    reference's RetinaFace.call."""
    g = golden('retinaface_call.npz')
    heads = [g[f'head{i}'] for i in range(9)]
    H, W = g['images'].shape[1:3]
    counts, _, rows, _ = run_native(heads, H, W)
    for n in range(len(counts)):
        k = counts[n]
        np.testing.assert_array_equal(rows[n, :k, 0], g[f'score{n}'])
        np.testing.assert_allclose(rows[n, :k, 1:5], g[f'bbox{n}'], rtol=0, atol=2e-4)
        np.testing.assert_allclose(rows[n, :k, 5:15].reshape(k, 5, 2), g[f'landmarks{n}'],
                                   rtol=0, atol=2e-4)


@pytest.mark.parametrize('shape', [(1, 416, 416), (3, 160, 232), (2, 75, 109), (4, 416, 739)])
def test_decode_nms_random(native, shape):
    N, H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    counts, cands = check_against_oracle(random_heads(rng, N, H, W), H, W)
    assert cands.min() > 0 and counts.min() > 0


def test_decode_nms_edge_cases(native):
    rng = np.random.default_rng(9)
    H, W = 96, 128
    heads = random_heads(rng, 2, H, W)
    # image 0: nothing passes; image 1 unchanged
    for i in (0, 3, 6):
        heads[i][0, 2:] = 0.1
        heads[i][0, :2] = 0.9
    counts, cands = check_against_oracle(heads, H, W)
    assert counts[0] == 0 and cands[0] == 0 and counts[1] > 0
    # everything passes (12k+ candidates on a full frame: the global-memory sort path)
    H, W = 416, 739
    heads = random_heads(rng, 1, H, W, frac=0.999)
    counts, cands = check_against_oracle(heads, H, W)
    assert cands[0] > 12000
    # threshold exactly equal to a score keeps it (>=)
    heads = random_heads(rng, 1, 64, 64)
    heads[0][0, 2, 0, 0] = 0.5
    check_against_oracle(heads, 64, 64)
    # heavy overlap: all deltas zero -> boxes are the anchors themselves
    heads = random_heads(rng, 1, 128, 128, frac=0.5)
    for i in (1, 4, 7):
        heads[i][:] = 0
    check_against_oracle(heads, 128, 128)


def test_full_size_nms_idempotent(native):
    """BASELINE size (32 x 416x739): survivors are mutually non-overlapping and
    re-running NMS on the survivors keeps all of them (size-independent
    properties, no oracle)."""
    rng = np.random.default_rng(4)
    H, W = 416, 739
    heads = random_heads(rng, 32, H, W, frac=0.01)
    counts, cands, rows, idx = run_native(heads, H, W)
    assert (cands > 50).all()
    for n in range(0, 32, 5):
        k = counts[n]
        sc = rows[n, :k, 0]
        assert (np.diff(sc) <= 0).all()                       # score-descending
        keep = detect.nms(rows[n, :k, 1:5], 0.4)
        assert len(keep) == k                                 # idempotent
        assert len(np.unique(idx[n, :k])) == k


# ------------------------------------------------------------------ pose parse

def native_parse(pafs, heats, scale):
    from terran_b200.pose.openpose.wrapper import parse_device, unpack_poses
    out = parse_device(torch.from_numpy(np.stack(pafs)).cuda(),
                       torch.from_numpy(np.stack(heats)).cuda(), scale)
    return unpack_poses(*out[:4])


def test_pose_parse_golden(native, golden):
    """The 24 scenes the reference itself parsed (fixture): identical keypoints."""
    g = golden('openpose_parse.npz')
    scale = 184 / 720
    scenes = [pose.synthetic_scene(int(s)) for s in g['seeds']]
    got = native_parse([s[0] for s in scenes], [s[1] for s in scenes], scale)
    for k, humans in enumerate(got):
        assert len(humans) == len(g[f'score{k}']), k
        if humans:
            np.testing.assert_array_equal(np.stack([h['keypoints'] for h in humans]), g[f'kp{k}'])
            np.testing.assert_allclose([h['score'] for h in humans], g[f'score{k}'], atol=1e-5)


def test_pose_parse_matches_oracle_exactly(native):
    """Fresh scenes (other seeds, crowded, other map sizes): keypoints AND the
    f64 scores bit-equal to the oracle (same fp32 operation order)."""
    for (h, w), seeds, people in (((23, 40), range(40, 52), None), ((23, 40), range(3), 14),
                                  ((17, 29), range(60, 64), None), ((48, 85), range(70, 72), 6)):
        scenes = [pose.synthetic_scene(s, h=h, w=w, people=people) for s in seeds]
        scale = 8 * h / 720
        got = native_parse([s[0] for s in scenes], [s[1] for s in scenes], scale)
        for (paf, heat), humans in zip(scenes, got):
            ref = pose.parse_frame(paf, heat, scale)
            assert len(humans) == len(ref)
            for a, b in zip(humans, ref):
                np.testing.assert_array_equal(a['keypoints'], b['keypoints'])
                assert a['score'] == b['score']


def test_pose_parse_edge_cases(native):
    z_paf, z_heat = np.zeros((38, 23, 40), np.float32), np.zeros((19, 23, 40), np.float32)
    one = z_heat.copy()
    one[0, 10, 10] = 1.0
    paf, heat = pose.synthetic_scene(5)
    got = native_parse([z_paf, z_paf, paf], [z_heat, one, heat], 0.25)
    assert got[0] == [] and got[1] == []
    assert len(got[2]) == len(pose.parse_frame(paf, heat, 0.25))


def test_pose_parse_capacity_overflow_truncates_one_frame(native):
    """More peaks than the per-part capacity on ONE frame: a RuntimeWarning, a truncated
    result for that frame, and the other frame of the batch exactly as if parsed alone (the
    reference has no capacities and never fails a batch: wrapper.py:226-483)."""
    from terran_b200 import _native as nat
    h, w = 46, 80
    paf, heat = pose.synthetic_scene(7, h=h, w=w, people=3)
    crowded = np.zeros_like(heat)
    yy, xx = np.mgrid[0:h, 0:w]
    crowded[0] = ((yy + xx) % 2 == 0).astype(np.float32)          # ~1800 isolated maxima of part 0
    assert crowded[0].sum() > nat.TR_PEAK_CAP
    scale = 8 * h / 720
    alone = native_parse([paf], [heat], scale)[0]
    with pytest.warns(RuntimeWarning, match='capacity exceeded on frame'):
        got = native_parse([np.zeros_like(paf), paf], [crowded, heat], scale)
    assert len(got) == 2 and len(got[1]) == len(alone) > 0
    for a, b in zip(got[1], alone):
        np.testing.assert_array_equal(a['keypoints'], b['keypoints'])
        assert a['score'] == b['score']


def test_pose_peaks_full_size(native):
    """BASELINE size (16 frames of 23x40 maps): every reported keypoint is a
    4-neighbour local maximum >= 0.1 of the up-sampled heat map."""
    scenes = [pose.synthetic_scene(100 + i) for i in range(16)]
    scale = 184 / 720
    got = native_parse([s[0] for s in scenes], [s[1] for s in scenes], scale)
    assert sum(len(g) for g in got) > 30
    for (paf, heat), humans in list(zip(scenes, got))[::5]:
        up = pose.bicubic_up8(heat)
        locs, _ = pose.find_peaks(up)
        allowed = [set((int(int(x) / scale), int(int(y) / scale)) for y, x in l) for l in locs]
        for hmn in humans:
            for j in range(18):
                x, y, present = hmn['keypoints'][j]
                if present:
                    assert (x, y) in allowed[j]
