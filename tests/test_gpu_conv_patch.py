"""GPU: the resident-patch tcgen05 kernel (csrc/conv_patch.cu, ``tr_conv2d(use_tc=3)``) against
an fp64 reference on the same fp16 operands — every tile orientation, ragged borders, all
epilogues, the BASELINE layer shapes."""
import os

import numpy as np
import pytest
import torch

from tests.gpu_util import conv2d_native, conv2d_reference, describe_mismatch

pytestmark = pytest.mark.gpu


def run_case(nat, N, H, W, cin, cout, k, *, act=1, res=False, out_f32=False, seed=0, env=None,
             repeat=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn((N, H, W, cin), generator=g) * 0.5).half().cuda()
    w = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    slope = torch.rand(cout, generator=g) * 0.3 if act == 2 else None
    r = (torch.randn((N, H, W, cout), generator=g) * 0.5).half().cuda() if res else None
    old = {}
    for key, val in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = str(val)
    try:
        out, _ = conv2d_native(nat, x, w, scale, shift, act=act, slope=slope, res=r,
                               out_f32=out_f32, use_tc=3, repeat=repeat)
    finally:
        for key, val in old.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
    ref = conv2d_reference(x, w, scale, shift, act=act, slope=slope, res=r)
    tol = 2e-3 if out_f32 else 4e-3
    err = (out.detach().cpu().double() - ref).abs().max().item()
    assert err < tol, describe_mismatch(out, ref, tol)
    return out


@pytest.mark.parametrize('shape', [
    (2, 23, 40, 128, 128, 7),       # OpenPose Mconv2-5 (both chunks, 49 taps)
    (1, 23, 40, 192, 256, 7),       # OpenPose Mconv1 (3 chunks, 2 cout tiles)
    (2, 23, 40, 256, 128, 3),       # conv4_4_CPM
    (1, 23, 40, 128, 512, 1),       # conv5_4_CPM (1x1: the patch is the tile, 8-pixel rows)
    (2, 46, 81, 128, 256, 3),       # conv3_1 (ragged on both axes)
    (3, 13, 17, 64, 128, 3),        # one chunk, tiny map
    (5, 7, 7, 128, 128, 3),         # ArcFace 7x7 maps
    (2, 9, 30, 64, 128, 5),         # 5x5
    (2, 23, 40, 64, 64, 3),         # 64 filters: half of a 128-row tile (TMA zero-fills / clips)
    (1, 30, 50, 128, 64, 3),
])
@pytest.mark.parametrize('dual', [0, 2], ids=['one-cta-per-sm', 'two-ctas-per-sm'])
def test_patch_conv_matches_reference(native, shape, dual):
    run_case(native, *shape, env={'TRB_PT_DUAL': dual})


@pytest.mark.parametrize('axis', [0, 1])
@pytest.mark.parametrize('R', [2, 6, 24, 32])
def test_patch_conv_tile_geometries(native, axis, R):
    run_case(native, 2, 23, 40, 128, 128, 3, env={'TRB_PT_AXIS': axis, 'TRB_PT_R': R})
    run_case(native, 1, 29, 21, 64, 128, 7, env={'TRB_PT_AXIS': axis, 'TRB_PT_R': R}, seed=1)


@pytest.mark.parametrize('sub,stages', [(1, 2), (2, 3), (3, 12), (3, 2)])
def test_patch_conv_ring_shapes(native, sub, stages):
    run_case(native, 2, 23, 40, 128, 128, 7, env={'TRB_PT_SUB': sub, 'TRB_PT_STAGES': stages})


def test_patch_conv_epilogues(native):
    run_case(native, 2, 14, 14, 128, 128, 3, act=2)                   # PReLU
    run_case(native, 2, 14, 14, 128, 128, 3, act=0, res=True)         # residual add
    run_case(native, 2, 14, 14, 128, 256, 3, act=2, res=True)
    run_case(native, 2, 23, 40, 128, 128, 1, act=0, out_f32=True)     # fp32 output
    run_case(native, 2, 23, 40, 128, 40, 1, act=0)                    # cout 40 of a 128 tile... padded filters
    run_case(native, 2, 28, 28, 64, 64, 3, act=2, res=True)           # 64 filters + residual


def test_patch_conv_many_tiles_per_cta(native):
    """More tiles than SMs: TMEM double buffering, patch buffer and ring phases wrap."""
    run_case(native, 40, 23, 40, 128, 128, 3)
    run_case(native, 64, 14, 14, 256, 256, 3, act=2, res=True)


@pytest.mark.parametrize('shape', [
    (32, 23, 40, 128, 128, 7),      # 160 tiles on 148 SMs: every CTA hands a partial tile over
    (2, 23, 40, 128, 128, 7),       # 10 tiles: split-K, each tile summed from ~15 CTAs
    (5, 23, 40, 192, 256, 3),       # three chunks, two cout tiles
    (40, 23, 40, 128, 128, 3),      # 200 tiles, short K
    (3, 9, 10, 64, 128, 1),         # one ring iteration per tile: nothing to split
])
@pytest.mark.parametrize('dual', [0, 2], ids=['one-cta-per-sm', 'two-ctas-per-sm'])
def test_patch_conv_stream_k(native, shape, dual):
    """Stream-K (TRB_PT_SK=2: whenever possible) against the reference, launched several times in
    a row (the hand-over flags re-arm themselves) and bit-identical to a second run."""
    a = run_case(native, *shape, env={'TRB_PT_SK': 2, 'TRB_PT_DUAL': dual}, repeat=3)
    b = run_case(native, *shape, env={'TRB_PT_SK': 2, 'TRB_PT_DUAL': dual}, repeat=1)
    assert torch.equal(a, b)


@pytest.mark.parametrize('dual', [0, 2], ids=['one-cta-per-sm', 'two-ctas-per-sm'])
def test_patch_conv_stream_k_epilogues(native, dual):
    run_case(native, 20, 14, 14, 128, 128, 3, act=2, res=True, env={'TRB_PT_SK': 2, 'TRB_PT_DUAL': dual}, repeat=2)
    run_case(native, 3, 23, 40, 128, 128, 3, act=0, out_f32=True, env={'TRB_PT_SK': 2, 'TRB_PT_DUAL': dual}, repeat=2)
    run_case(native, 40, 28, 28, 128, 128, 3, act=2, res=True, env={'TRB_PT_DUAL': dual})       # many tiles per CTA
    run_case(native, 32, 23, 40, 128, 128, 7, env={'TRB_PT_SK': 2, 'TRB_PT_SUB': 1, 'TRB_PT_STAGES': 3})


@pytest.mark.parametrize('stack', [0, 1], ids=['one-image-per-tile', 'stacked-images'])
def test_patch_conv_stacked_images(native, stack):
    """Small maps: several images per tile with garbage lines between them (ArcFace 14x14 and
    7x7 stages, arcface/model.py:11-35) — odd batch sizes leave the last tile short of images;
    residual and PReLU go through the line -> (image, row) mapping of the epilogue."""
    env = {'TRB_PT_STACK': stack}
    run_case(native, 5, 14, 14, 256, 256, 3, act=0, res=True, env=env)
    run_case(native, 7, 7, 7, 128, 256, 3, act=2, env=env, seed=2)
    run_case(native, 3, 14, 9, 64, 128, 5, env=env, seed=3)           # 5x5: bstride 14 + 4
    run_case(native, 2, 6, 28, 128, 128, 3, act=2, res=True, env=env, seed=4)
