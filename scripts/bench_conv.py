"""Micro-benchmark of single convolutions through tr_conv2d (CUDA events inside
the library, `repeat` back-to-back launches).  TRB_TC_DEBUG=1/2 isolates the
TMA and MMA halves of the pipeline (timing only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from terran_b200 import _native as nat
from tests.gpu_util import conv2d_native

CASES = [  # N, H, W, cin, cout, k, stride, tag
    (32, 23, 40, 128, 128, 7, 1, 'openpose 7x7 128->128'),
    (32, 23, 40, 192, 128, 7, 1, 'openpose 7x7 185->128'),
    (32, 23, 40, 512, 512, 3, 1, 'vgg 3x3 512->512'),
    (32, 46, 81, 256, 256, 3, 1, 'vgg 3x3 256->256'),
    (32, 184, 327, 64, 64, 3, 1, 'vgg 3x3 64->64 full res'),
    (32, 23, 40, 128, 128, 1, 1, 'openpose 1x1 128->128'),
    (32, 23, 40, 128, 512, 1, 1, 'openpose 1x1 128->512'),
    (32, 23, 40, 512, 38, 1, 1, 'openpose 1x1 512->38'),
    (32, 92, 163, 64, 128, 3, 1, 'vgg 3x3 64->128 @92x163'),
    (32, 92, 163, 128, 128, 3, 1, 'vgg 3x3 128->128 @92x163'),
    (256, 56, 56, 64, 64, 3, 1, 'arcface 3x3 64 @56'),
    (256, 28, 28, 128, 128, 3, 1, 'arcface 3x3 128 @28'),
    (256, 14, 14, 256, 256, 3, 1, 'arcface 3x3 256 @14'),
    (256, 7, 7, 512, 512, 3, 1, 'arcface 3x3 512 @7'),
    (32, 52, 93, 64, 64, 3, 1, 'retina 3x3 64->64 @52x93'),
    (32, 26, 47, 128, 128, 1, 1, 'retina 1x1 128->128 @26x47'),
]
SEL = [a for a in sys.argv[1:] if not a.startswith('-')]
REPEAT = 3 if '--short' in sys.argv else 20
ENGINE = 3 if '--patch' in sys.argv else 1
for (N, H, W, cin, cout, k, stride, tag) in CASES:
    if SEL and not any(s in tag for s in SEL):
        continue
    if ENGINE == 3 and (stride != 1 or cin % 64 or cout % 128):
        continue
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((N, H, W, cin), generator=g) * 0.5).half().cuda()
    w = torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
    out, ms = conv2d_native(nat, x, w, torch.ones(cout), torch.zeros(cout), stride=stride, act=1,
                            use_tc=ENGINE, repeat=REPEAT)
    Ho, Wo = out.shape[1:3]
    fl = 2.0 * N * Ho * Wo * cout * cin * k * k
    print(f'{tag:28s} {ms * 1e3:9.1f} us {fl / ms / 1e9:8.1f} TFLOP/s', flush=True)
