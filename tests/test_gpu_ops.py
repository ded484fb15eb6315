"""GPU: single-op parity through the C ABI — the tcgen05 implicit-GEMM conv
and the CUDA-core kernels against an fp64 reference on identical fp16 operands."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.gpu_util import conv2d_native, conv2d_reference, describe_mismatch, \
    sepconv2d_native, sepconv2d_reference

pytestmark = pytest.mark.gpu

# (N, H, W, cin, cout, k, stride, act, residual, tag)
CONV_CASES = [
    (1, 8, 16, 64, 64, 1, 1, 0, None, 'gemm-64x64'),
    (1, 8, 16, 64, 16, 1, 1, 0, None, 'gemm-n16'),
    (2, 9, 13, 64, 128, 1, 1, 1, None, 'gemm-ragged'),
    (1, 8, 16, 128, 64, 1, 1, 0, None, 'gemm-2chunks'),
    (1, 8, 16, 32, 32, 1, 1, 0, None, 'gemm-sw64'),
    (1, 8, 16, 16, 16, 1, 1, 0, None, 'gemm-sw32'),
    (1, 12, 20, 64, 64, 3, 1, 1, None, 'conv3-64'),
    (2, 13, 24, 64, 32, 3, 1, 1, None, 'conv3-ctx'),
    (1, 13, 24, 16, 16, 3, 1, 1, None, 'conv3-sw32'),
    (1, 23, 40, 128, 128, 7, 1, 1, None, 'conv7-openpose'),
    (2, 23, 40, 192, 128, 7, 1, 1, None, 'conv7-cat192'),
    (1, 23, 40, 128, 38, 1, 1, 0, None, 'head-38'),
    (3, 14, 14, 256, 256, 3, 1, 2, 'same', 'arcface-unit'),
    (2, 28, 28, 128, 128, 3, 2, 0, 'same', 'arcface-s2'),
    (2, 28, 28, 64, 128, 1, 2, 0, None, 'arcface-shortcut'),
    (5, 7, 7, 512, 512, 3, 1, 2, None, 'arcface-7x7'),
    (1, 26, 47, 128, 64, 1, 1, 1, 'up2', 'fpn-lateral'),
    (1, 46, 81, 256, 512, 3, 1, 1, None, 'vgg-512'),
    (130, 1, 1, 1024, 512, 1, 1, 0, None, 'fc-like'),
]


def make_case(case, seed=0):
    N, H, W, cin, cout, k, stride, act, res_kind, _ = case
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn((N, H, W, cin), generator=g) * 0.5).half().cuda()
    w = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    scale = torch.empty(cout).uniform_(0.5, 1.5, generator=g)
    shift = torch.randn(cout, generator=g) * 0.1
    slope = torch.empty(cout).uniform_(0.1, 0.4, generator=g)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res, up2 = None, False
    cstore = (cout + 7) // 8 * 8
    if res_kind == 'same':
        res = (torch.randn((N, Ho, Wo, cstore), generator=g) * 0.5).half().cuda()
    elif res_kind == 'up2':
        res = (torch.randn((N, (Ho + 1) // 2, (Wo + 1) // 2, cstore), generator=g) * 0.5).half().cuda()
        up2 = True
    return x, w, scale, shift, slope, res, up2


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[-1] for c in CONV_CASES])
@pytest.mark.parametrize('use_tc', [False, True], ids=['direct', 'tcgen05'])
def test_conv_matches_reference(native, case, use_tc):
    N, H, W, cin, cout, k, stride, act, res_kind, tag = case
    x, w, scale, shift, slope, res, up2 = make_case(case)
    out, _ = conv2d_native(native, x, w, scale, shift, stride=stride, act=act, slope=slope,
                           res=res, res_up2=up2, use_tc=use_tc)
    ref = conv2d_reference(x, w, scale, shift, stride=stride, act=act, slope=slope, res=res,
                           res_up2=up2)
    tol = 2e-3 * max(1.0, float(ref.abs().max()))      # fp16 output rounding
    err = (out.cpu().double() - ref).abs().max()
    assert torch.isfinite(out.float()).all() and err <= tol, describe_mismatch(out, ref, tol)


# Stream-K (the CTAs split the k-iterations of all tiles evenly; split tiles are completed
# through an fp32 partial in global memory): shapes with more tiles than SMs.
STREAMK_CASES = [
    (24, 23, 40, 64, 64, 3, 1, 1, None, 'sk-conv3'),
    (24, 23, 40, 512, 256, 1, 1, 2, 'same', 'sk-gemm-res'),
    (24, 46, 80, 128, 128, 3, 2, 0, None, 'sk-stride2'),
    (12, 23, 40, 128, 512, 3, 1, 1, None, 'sk-2ntiles'),
    (20, 23, 40, 128, 128, 7, 1, 1, None, 'sk-conv7'),
]


@pytest.mark.parametrize('case', STREAMK_CASES, ids=[c[-1] for c in STREAMK_CASES])
def test_conv_stream_k(native, case, monkeypatch):
    N, H, W, cin, cout, k, stride, act, res_kind, tag = case
    x, w, scale, shift, slope, res, up2 = make_case(case)
    kw = dict(stride=stride, act=act, slope=slope, res=res, res_up2=up2)
    monkeypatch.setenv('TRB_TC_HALO', '0')
    monkeypatch.setenv('TRB_TC_SWAP', '0')
    monkeypatch.setenv('TRB_TC_SK', '0')
    whole, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, out_f32=True, **kw)
    monkeypatch.setenv('TRB_TC_SK', '2')
    split, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, out_f32=True, **kw)
    again, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, out_f32=True, **kw)
    ref = conv2d_reference(x, w, scale, shift, **kw)
    tol = 1e-4 * max(1.0, float(ref.abs().max()))
    assert (split.cpu().double() - ref).abs().max() <= tol, describe_mismatch(split, ref, tol)
    assert (whole.cpu().double() - ref).abs().max() <= tol
    # the split really happened (a different fp32 summation order shows in the last bit of
    # some outputs) and is deterministic
    assert not torch.equal(split, whole)
    assert torch.equal(split, again)


# Warp-level mma.sync kernel (conv_mma.cu): RetinaFace's refiner / context / head layers.
# (N, H, W, cin, cout, k, stride, act, residual, tag) — odd sizes exercise the strip masks.
MMA_CASES = [
    (2, 13, 24, 256, 64, 1, 1, 1, None, 'lateral-s32'),
    (2, 26, 47, 128, 64, 1, 1, 1, 'up2', 'lateral-s16-up2'),
    (1, 52, 93, 64, 64, 1, 1, 1, 'up2', 'lateral-s8-up2'),
    (2, 26, 47, 64, 64, 3, 1, 1, None, 'aggr'),
    (3, 13, 24, 64, 32, 3, 1, 1, None, 'ctx-3x3'),
    (1, 27, 45, 64, 16, 3, 1, 1, None, 'ctx-reducer'),
    (2, 13, 24, 16, 16, 3, 1, 1, None, 'ctx-16'),
    (5, 7, 9, 64, 64, 3, 1, 0, 'same', 'res-same'),
    (1, 1, 1, 64, 32, 1, 1, 0, None, 'one-pixel'),
]


@pytest.mark.parametrize('case', MMA_CASES, ids=[c[-1] for c in MMA_CASES])
def test_conv_mma_matches_reference(native, case):
    N, H, W, cin, cout, k, stride, act, res_kind, tag = case
    x, w, scale, shift, slope, res, up2 = make_case(case)
    out, _ = conv2d_native(native, x, w, scale, shift, stride=stride, act=act, res=res,
                           res_up2=up2, use_tc=2)
    ref = conv2d_reference(x, w, scale, shift, stride=stride, act=act, res=res, res_up2=up2)
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    err = (out.cpu().double() - ref).abs().max()
    assert torch.isfinite(out.float()).all() and err <= tol, describe_mismatch(out, ref, tol)


def test_conv_mma_fp32_head_into_channel_slice(native):
    """Fused head conv: fp32 output, no activation (retinaface/model.py:248-316)."""
    case = (2, 13, 24, 64, 32, 1, 1, 0, None, 'head')
    x, w, scale, shift, slope, res, up2 = make_case(case, seed=3)
    out, _ = conv2d_native(native, x, w, scale, shift, use_tc=2, out_f32=True)
    ref = conv2d_reference(x, w, scale, shift)
    assert out.dtype == torch.float32
    assert (out.cpu().double() - ref).abs().max() < 1e-4, describe_mismatch(out, ref, 1e-4)


# Fused depthwise 3x3 + BN + ReLU -> 1x1 + BN + ReLU: every (channels, stride) pair of the
# mobilenet-0.25 backbone (retinaface/model.py:53-112), ragged map sizes.
SEP_CASES = [
    (2, 21, 70, 8, 16, 1), (1, 20, 67, 16, 32, 2), (2, 11, 19, 32, 32, 1), (2, 21, 35, 32, 64, 2),
    (1, 13, 24, 64, 64, 1), (3, 13, 23, 64, 128, 2), (2, 9, 17, 128, 128, 1), (2, 13, 24, 128, 256, 2),
    (2, 7, 12, 256, 256, 1), (1, 1, 1, 8, 16, 1), (33, 2, 3, 128, 128, 1),
]


def make_sep_case(case, seed=0):
    N, H, W, cin, cout, stride = case
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn((N, H, W, cin), generator=g) * 0.5).abs().half().cuda()
    dw_w = torch.randn((cin, 1, 3, 3), generator=g) / 3.0
    dw_scale = torch.empty(cin).uniform_(0.5, 1.5, generator=g)
    dw_shift = torch.randn(cin, generator=g) * 0.1
    w = torch.randn((cout, cin, 1, 1), generator=g) / cin ** 0.5
    scale = torch.empty(cout).uniform_(0.5, 1.5, generator=g)
    shift = torch.randn(cout, generator=g) * 0.1
    return x, dw_w, dw_scale, dw_shift, w, scale, shift


@pytest.mark.parametrize('case', SEP_CASES, ids=['x'.join(map(str, c)) for c in SEP_CASES])
@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'unfused'])
def test_sepconv_matches_reference(native, case, fused):
    args = make_sep_case(case)
    out, _ = sepconv2d_native(native, *args, stride=case[5], fused=fused)
    ref = sepconv2d_reference(*args, stride=case[5])
    # the depthwise result is rounded to fp16 before the 1x1: one half-ulp flip of an
    # operand moves the output by ~1e-3 of its scale
    tol = 4e-3 * max(1.0, float(ref.abs().max()))
    err = (out.cpu().double() - ref).abs().max()
    assert torch.isfinite(out.float()).all() and err <= tol, describe_mismatch(out, ref, tol)


# Swap mode (filters on the UMMA M axis, a band of full-width pixel rows on N <= 256; TMEM holds
# the transposed tile): forced wherever the layer allows it, ragged last bands, two filter
# tiles, PReLU, and combined with stream-K.
SWAP_CASES = [
    (1, 23, 40, 128, 128, 7, 1, 1, None, 'swap-conv7-openpose'),
    (2, 23, 40, 192, 128, 7, 1, 1, None, 'swap-conv7-cat192'),
    (3, 13, 24, 64, 128, 3, 1, 2, None, 'swap-ragged-band-prelu'),
    (1, 23, 40, 64, 256, 3, 1, 0, None, 'swap-two-filter-tiles'),
    (2, 9, 16, 128, 120, 1, 1, 1, None, 'swap-cout120'),
    (2, 46, 81, 64, 128, 5, 1, 1, None, 'swap-5x5-wide-map'),
]


@pytest.mark.parametrize('variant', ['2', '3'], ids=['patch', 'band'])
@pytest.mark.parametrize('case', SWAP_CASES, ids=[c[-1] for c in SWAP_CASES])
def test_conv_swap_mode(native, case, variant, monkeypatch):
    N, H, W, cin, cout, k, stride, act, res_kind, tag = case
    x, w, scale, shift, slope, res, up2 = make_case(case)
    kw = dict(stride=stride, act=act, slope=slope)
    monkeypatch.setenv('TRB_TC_HALO', '0')
    monkeypatch.setenv('TRB_TC_SWAP', '0')
    plain, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, **kw)
    # 2: filters x columns of the resident halo patch where the filter allows (k >= 3), else
    # 3: filters x a band of full-width rows
    monkeypatch.setenv('TRB_TC_SWAP', variant)
    out, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, **kw)
    ref = conv2d_reference(x, w, scale, shift, **kw)
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    err = (out.cpu().double() - ref).abs().max()
    assert torch.isfinite(out.float()).all() and err <= tol, describe_mismatch(out, ref, tol)
    # same fp16 operands and fp32 accumulation over the same k order: the two tilings agree
    # to fp16 output rounding
    assert (out.float() - plain.float()).abs().max() <= tol


def test_conv_swap_mode_stream_k(native, monkeypatch):
    case = (40, 23, 40, 128, 128, 3, 1, 1, None, 'swap-sk')
    x, w, scale, shift, slope, res, up2 = make_case(case)
    monkeypatch.setenv('TRB_TC_HALO', '0')
    monkeypatch.setenv('TRB_TC_SWAP', '3')
    monkeypatch.setenv('TRB_TC_SK', '0')
    whole, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, act=1)
    monkeypatch.setenv('TRB_TC_SK', '2')
    split, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, act=1)
    ref = conv2d_reference(x, w, scale, shift, act=1)
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    assert (split.cpu().double() - ref).abs().max() <= tol, describe_mismatch(split, ref, tol)
    assert (whole.cpu().double() - ref).abs().max() <= tol, describe_mismatch(whole, ref, tol)


def test_conv_tc_fp32_output(native):
    case = (2, 13, 24, 64, 32, 1, 1, 0, None, 'head')
    x, w, scale, shift, slope, res, up2 = make_case(case, seed=3)
    out, _ = conv2d_native(native, x, w, scale, shift, use_tc=True, out_f32=True)
    ref = conv2d_reference(x, w, scale, shift)
    assert out.dtype == torch.float32
    assert (out.cpu().double() - ref).abs().max() < 1e-4, describe_mismatch(out, ref, 1e-4)


def test_resize_matches_cv2(native):
    """The resize kernel is a bit-exact restatement of cv2.resize(INTER_LINEAR)
    on uint8 (reference host resize, face/detection/__init__.py:15-57)."""
    import cv2
    from terran_b200.frames import resize_short_side
    rng = np.random.default_rng(0)
    for (H, W), short in (((1080, 1920), 416), ((720, 1280), 184), ((640, 640), 416),
                          ((1080, 1920), 184), ((333, 517), 416), ((97, 61), 184),
                          ((2160, 3840), 184), ((300, 401), 184), ((277, 1003), 97)):
        frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        got, scale = resize_short_side(torch.from_numpy(frames).cuda(), short)
        size = (int(W * scale), int(H * scale))
        for i in range(2):
            want = cv2.resize(frames[i], size, interpolation=cv2.INTER_LINEAR)
            np.testing.assert_array_equal(got[i].cpu().numpy(), want, err_msg=f'{H}x{W}->{short}')


def test_l2_normalize(native):
    native.init(0)
    rng = np.random.default_rng(1)
    x = rng.normal(size=(37, 512)).astype(np.float32)
    x[5] = 0                      # zero row divides by 1 (sklearn semantics)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    native.check(native.lib().tr_l2_normalize(C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()),
                                              37, 512, native.current_stream_ptr()))
    norm = np.sqrt((x ** 2).sum(1, keepdims=True))
    norm[norm == 0] = 1
    np.testing.assert_allclose(out.cpu().numpy(), x / norm, rtol=0, atol=1e-6)


def test_bicubic_table_matches_oracle(native):
    from oracle import pose
    tab = (C.c_float * 32)()
    native.lib().tr_bicubic_table(tab)
    np.testing.assert_array_equal(np.array(tab, np.float32).reshape(8, 4), pose.bicubic_table())


def test_device_similarity_matches_host_on_10k_faces(native):
    """``tr_face_similarity`` (closed form on the device, straight from detection rows) against the
    host ``similarity_coefficients`` — itself pinned against the oracle's SVD Umeyama — on
    10 000 random faces: landmarks are mapped back with value / scale, round half to even, as
    ``Detection.resize_out`` does (detection/__init__.py:59-84)."""
    import ctypes as C
    from terran_b200 import _native as nat
    from terran_b200.face.recognition.arcface.wrapper import LANDMARK_TEMPLATE, similarity_coefficients
    rng = np.random.default_rng(5)
    N, max_det, scale = 40, 256, 416 / 1080
    counts = rng.integers(200, max_det + 1, N).astype(np.int32)
    det = np.zeros((N, max_det, 16), np.float32)
    th = rng.uniform(-np.pi, np.pi, (N, max_det))
    sc = rng.uniform(0.3, 4.0, (N, max_det))
    R = np.stack([np.stack([np.cos(th), -np.sin(th)], -1), np.stack([np.sin(th), np.cos(th)], -1)], -2)
    lm = sc[..., None, None] * np.einsum('nfij,kj->nfki', R, LANDMARK_TEMPLATE.astype(np.float64)) \
        + rng.uniform(0, 700, (N, max_det, 1, 2)) + rng.normal(0, 0.7, (N, max_det, 5, 2))
    det[..., 5:15] = lm.reshape(N, max_det, 10)
    F = int(counts.sum())
    assert F >= 8000
    dev = torch.device('cuda')
    d_det, d_cnt = torch.from_numpy(det).to(dev), torch.from_numpy(counts).to(dev)
    coef = torch.empty((F, 6), dtype=torch.float64, device=dev)
    idx = torch.empty(F, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.check(nat.lib().tr_face_similarity(
        C.c_void_p(d_det.data_ptr()), C.c_void_p(d_cnt.data_ptr()), N, max_det, C.c_float(scale), F,
        C.c_void_p(coef.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(total.data_ptr()),
        nat.current_stream_ptr()))
    torch.cuda.synchronize()
    assert int(total.item()) == F
    want_idx = np.repeat(np.arange(N), counts)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    rows = np.concatenate([det[n, :counts[n], 5:15] for n in range(N)])
    pix = np.around(rows / np.float32(scale)).astype(np.int32).reshape(F, 5, 2)     # resize_out rounding
    want = similarity_coefficients(pix)
    got = coef.cpu().numpy()
    rel = np.abs(got - want) / np.maximum(1e-3, np.abs(want))
    assert rel.max() < 1e-9, rel.max()
