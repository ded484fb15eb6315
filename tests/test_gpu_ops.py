"""GPU: single-op parity through the C ABI — the tcgen05 implicit-GEMM conv
and the CUDA-core kernels against an fp64 reference on identical fp16 operands."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.gpu_util import conv2d_native, conv2d_reference, describe_mismatch

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
                          ((1080, 1920), 184), ((333, 517), 416), ((97, 61), 184)):
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
