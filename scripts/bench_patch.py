"""Micro-benchmark of the resident-patch kernel (tr_conv2d use_tc=3): tile-width sweep and the
timing-only debug modes (TRB_PT_DEBUG: 1 no filter TMA, 2 no MMA, 4 no stores) on shapes whose
tile count is an exact multiple of the SM count, so that the time per filter block is visible.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from terran_b200 import _native as nat
from tests.gpu_util import conv2d_native

SMS = torch.cuda.get_device_properties(0).multi_processor_count
CLK = float(os.environ.get('CLK_GHZ', '1.965'))


def run(tag, N, H, W, cin, cout, k, env, engine=3, repeat=20):
    old = {}
    for key, val in env.items():
        old[key] = os.environ.get(key); os.environ[key] = str(val)
    try:
        g = torch.Generator().manual_seed(0)
        x = (torch.randn((N, H, W, cin), generator=g) * 0.5).half().cuda()
        w = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
        out, ms = conv2d_native(nat, x, w, torch.ones(cout), torch.zeros(cout), act=1, use_tc=engine, repeat=repeat)
    finally:
        for key, val in old.items():
            if val is None: os.environ.pop(key, None)
            else: os.environ[key] = val
    fl = 2.0 * N * H * W * cout * cin * k * k
    kblocks = k * k * (cin // 64)
    print(f'{tag:44s} {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TFLOP/s  env {env}', flush=True)
    return ms


MODE = sys.argv[1] if len(sys.argv) > 1 else 'sweep'
if MODE == 'fc':
    # ArcFace's 25088 -> 512 FC as a (1, 32, 8, 25088) map: split-K over all SMs, patch-buffer depth
    for nbuf in (2, 4, 8):
        for R in (8, 16):
            run(f'arcface fc 25088->512 x256 R={R} nbuf={nbuf}', 1, 32, 8, 25088, 512, 1,
                {'TRB_PT_R': R, 'TRB_PT_NBUF': nbuf, 'TRB_PT_AXIS': 0})
    run('trace arcface fc nbuf=8', 1, 32, 8, 25088, 512, 1, {'TRB_PT_R': 8, 'TRB_PT_AXIS': 0, 'TRB_PT_DEBUG': 32}, repeat=6)
    sys.exit(0)
if MODE == 'epi':
    # is the epilogue the bound of the short-K layers?  (64: no epilogue at all, 4: no stores, 2: no MMAs)
    for tag, shape in (('vgg 3x3 64->128 @92x163', (32, 92, 163, 64, 128, 3)),
                       ('vgg 3x3 128->128 @92x163', (32, 92, 163, 128, 128, 3)),
                       ('vgg 3x3 256->256 @46x81', (32, 46, 81, 256, 256, 3)),
                       ('arcface 3x3 128 @28', (256, 28, 28, 128, 128, 3))):
        for dbg in (0, 64, 4, 2, 66):
            run(f'{tag} dbg={dbg}', *shape, {'TRB_PT_SK': 0, 'TRB_PT_DEBUG': dbg})
    sys.exit(0)
if MODE == 'dual':
    for dual in (0, 2):
        run(f'openpose 7x7 128->128 dual={dual}', 32, 23, 40, 128, 128, 7, {'TRB_PT_DUAL': dual})
        run(f'openpose 7x7 185->256 dual={dual}', 32, 23, 40, 192, 256, 7, {'TRB_PT_DUAL': dual})
        run(f'vgg 3x3 512->512 @23x40 dual={dual}', 32, 23, 40, 512, 512, 3, {'TRB_PT_DUAL': dual})
        run(f'vgg 3x3 256->256 @46x81 dual={dual}', 32, 46, 81, 256, 256, 3, {'TRB_PT_DUAL': dual})
        run(f'vgg 3x3 128->128 @92x163 dual={dual}', 32, 92, 163, 128, 128, 3, {'TRB_PT_DUAL': dual})
        run(f'openpose 3x3 128->128 @23x40 dual={dual}', 32, 23, 40, 128, 128, 3, {'TRB_PT_DUAL': dual})
        run(f'openpose 1x1 128->512 @23x40 dual={dual}', 32, 23, 40, 128, 512, 1, {'TRB_PT_DUAL': dual})
        run(f'openpose 1x1 128->128 @23x40 dual={dual}', 32, 23, 40, 128, 128, 1, {'TRB_PT_DUAL': dual})
        run(f'arcface 3x3 128 @28 dual={dual}', 256, 28, 28, 128, 128, 3, {'TRB_PT_DUAL': dual})
        run(f'arcface 3x3 256 @14 dual={dual}', 256, 14, 14, 256, 256, 3, {'TRB_PT_DUAL': dual})
    run('trace openpose 7x7 128->128 dual=2', 32, 23, 40, 128, 128, 7, {'TRB_PT_DUAL': 2, 'TRB_PT_DEBUG': 32}, repeat=6)
    sys.exit(0)
if MODE == 'sk':
    for sk in (0, 2):
        run(f'openpose 7x7 128->128 sk={sk}', 32, 23, 40, 128, 128, 7, {'TRB_PT_SK': sk})
        run(f'openpose 7x7 185->256 sk={sk}', 32, 23, 40, 192, 256, 7, {'TRB_PT_SK': sk})
        run(f'vgg 3x3 512->512 @23x40 sk={sk}', 32, 23, 40, 512, 512, 3, {'TRB_PT_SK': sk})
        run(f'vgg 3x3 256->256 @46x81 sk={sk}', 32, 46, 81, 256, 256, 3, {'TRB_PT_SK': sk})
        run(f'vgg 3x3 128->128 @92x163 sk={sk}', 32, 92, 163, 128, 128, 3, {'TRB_PT_SK': sk})
        run(f'openpose 3x3 128->128 @23x40 sk={sk}', 32, 23, 40, 128, 128, 3, {'TRB_PT_SK': sk})
        run(f'openpose 1x1 128->512 @23x40 sk={sk}', 32, 23, 40, 128, 512, 1, {'TRB_PT_SK': sk})
        run(f'arcface 3x3 256 @14 sk={sk}', 256, 14, 14, 256, 256, 3, {'TRB_PT_SK': sk})
        run(f'arcface 3x3 512 @7 sk={sk}', 256, 7, 7, 512, 512, 3, {'TRB_PT_SK': sk})
        run(f'arcface 3x3 128 @28 sk={sk}', 256, 28, 28, 128, 128, 3, {'TRB_PT_SK': sk})
    run('trace openpose 7x7 128->128 sk=2', 32, 23, 40, 128, 128, 7, {'TRB_PT_SK': 2, 'TRB_PT_DEBUG': 32}, repeat=6)
    sys.exit(0)
if MODE == 'trace':
    run('trace 7x7 128->128 one tile/SM R=24 sub=2', 37, 24, 32, 128, 128, 7, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 24, 'TRB_PT_SUB': 2, 'TRB_PT_DEBUG': 32}, repeat=8)
    run('trace openpose 7x7 128->128 sub=2', 32, 23, 40, 128, 128, 7, {'TRB_PT_SUB': 2, 'TRB_PT_DEBUG': 32}, repeat=8)
    run('trace 3x3 256->256 @46x81', 32, 46, 81, 256, 256, 3, {'TRB_PT_SUB': 2, 'TRB_PT_DEBUG': 32}, repeat=6)
    sys.exit(0)
if MODE == 'issue':
    # what the issuing thread pays: handshake pieces switched off one by one (no filter TMA)
    for R in (24, 32):
        for sub in (1, 2):
            for dbg in (1,):
                ms = run(f'7x7 128->128 one tile/SM R={R} N={8*R} dbg={dbg} sub={sub}', 37, R, 32, 128, 128, 7,
                         {'TRB_PT_AXIS': 0, 'TRB_PT_R': R, 'TRB_PT_DEBUG': dbg, 'TRB_PT_SUB': sub})
                print(f'      -> {ms * 1e-3 * CLK * 1e9 / 98:8.0f} cycles per filter block')
    run('7x7 128->128 two tiles/SM R=24', 74, 24, 32, 128, 128, 7, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 24})
    run('7x7 128->128 two tiles/SM R=24 sub=2', 74, 24, 32, 128, 128, 7, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 24, 'TRB_PT_SUB': 2})
    run('7x7 128->128 four tiles/SM R=24 sub=2', 148, 24, 32, 128, 128, 7, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 24, 'TRB_PT_SUB': 2})
    sys.exit(0)
# 148 tiles exactly: W = 32 (4 strips of 8), H = R, N = 37 images; 7x7 128->128 = 98 filter blocks per tile
for R in (16, 24, 32):
    for dbg in (0, 1, 2, 3):
        for sub in (1, 2):
            ms = run(f'7x7 128->128 one tile/SM R={R} N={8*R} dbg={dbg} sub={sub}', 37, R, 32, 128, 128, 7,
                     {'TRB_PT_AXIS': 0, 'TRB_PT_R': R, 'TRB_PT_DEBUG': dbg, 'TRB_PT_SUB': sub})
            print(f'      -> {ms * 1e-3 * CLK * 1e9 / 98:8.0f} cycles per filter block at {CLK} GHz (MMA floor {4 * 8 * R // 2})')
# two tiles per SM (TMEM double buffering, patch re-load between tiles)
run('7x7 128->128 two tiles/SM R=32', 74, 32, 32, 128, 128, 7, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 32})
run('3x3 256->256 one tile/SM R=32', 37, 32, 32, 256, 128, 3, {'TRB_PT_AXIS': 0, 'TRB_PT_R': 32})
# the real layer, stage variations
for st in (3, 6):
    run(f'openpose 7x7 128->128 stages={st}', 32, 23, 40, 128, 128, 7, {'TRB_PT_STAGES': st})
run('openpose 7x7 128->128 plain kernel', 32, 23, 40, 128, 128, 7, {}, engine=1)
