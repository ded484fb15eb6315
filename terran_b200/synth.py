"""Synthetic, seeded checkpoints in the reference's ``state_dict`` key layout.

The real checkpoints are GitHub release downloads (reference
``terran/checkpoint.py:49-52,73-76,98-101``) and cannot be fetched offline, so
parity tests, ``smoke()`` and ``bench.py`` all run on weights generated here.
The key names and tensor shapes follow SURVEY.md Appendix C; they are verified
against the reference ``nn.Module``s with ``load_state_dict(strict=True)`` by
``oracle/make_golden.py``.

Weights are drawn He-style (``N(0, gain^2 / fan_in)``) rather than with the
``nn.Module`` default so that activations neither vanish nor explode through
the 56 / 100 / 92-conv stacks; BatchNorm running statistics are randomised so
that BN folding is actually exercised.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    return g


def _conv_w(g, cout, cin_per_group, k, gain=math.sqrt(2.0)):
    fan_in = cin_per_group * k * k
    return torch.randn(cout, cin_per_group, k, k, generator=g) * (
        gain / math.sqrt(fan_in)
    )


def _bn(sd, prefix, g, c, gamma=(0.8, 1.2), beta=0.1, mean=0.1, var=(0.5, 2.0)):
    sd[prefix + '.weight'] = torch.empty(c).uniform_(*gamma, generator=g)
    sd[prefix + '.bias'] = torch.randn(c, generator=g) * beta
    sd[prefix + '.running_mean'] = torch.randn(c, generator=g) * mean
    sd[prefix + '.running_var'] = torch.empty(c).uniform_(*var, generator=g)
    sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


# ---------------------------------------------------------------------------
# RetinaFace (mnet-0.25).  Layout: reference retinaface/model.py:53-316.
# ---------------------------------------------------------------------------

#: (in_c, out_c, stride) of the ConvSepBlocks of ``base.scales`` (model.py:74-91).
RETINAFACE_SCALES = (
    ((8, 16, 2), (16, 32, 1), (32, 32, 2), (32, 64, 1), (64, 64, 2)),
    ((64, 128, 1), (128, 128, 1), (128, 128, 1), (128, 128, 1), (128, 128, 1),
     (128, 128, 2)),
)


def retinaface_state_dict(seed=3, cls_gain=30.0, cls_bias=(-10.0, -5.5, -21.5)):
    """Seeded RetinaFace checkpoint.

    ``cls_gain`` scales the class-head weights and ``cls_bias`` (one value per
    stride 32, 16, 8) shifts the foreground logits so that foreground
    probabilities are well separated from the 0.5 threshold and only ~1 % of
    the anchors of a uniform-noise frame passes (SURVEY.md section 8(d):
    random init clusters every score at ~0.5).  The defaults were calibrated
    for seed 3 on ``default_rng`` uint8 noise frames.
    """
    g = _gen(seed)
    sd = {}

    def conv_bn(prefix_conv, prefix_bn, cin, cout, k, groups=1, bias=False):
        sd[prefix_conv + '.weight'] = _conv_w(g, cout, cin // groups, k)
        if bias:
            sd[prefix_conv + '.bias'] = torch.randn(cout, generator=g) * 0.05
        _bn(sd, prefix_bn, g, cout)

    # Input pixels are raw 0..255 (wrapper.py:144-146): scale the stem so the
    # first activation is O(1).
    sd['base.first_conv_block.0.weight'] = _conv_w(g, 8, 3, 3) / 128.0
    _bn(sd, 'base.first_conv_block.1', g, 8, mean=0.5)
    conv_bn('base.first_conv_block.3', 'base.first_conv_block.4', 8, 8, 3, groups=8)

    def sep_block(prefix, cin, cout):
        conv_bn(prefix + '.conv_block.0', prefix + '.conv_block.1', cin, cout, 1)
        conv_bn(prefix + '.sep_block.0', prefix + '.sep_block.1', cout, cout, 3,
                groups=cout)

    for si, blocks in enumerate(RETINAFACE_SCALES):
        for bi, (cin, cout, _stride) in enumerate(blocks):
            sep_block(f'base.scales.{si}.{bi}', cin, cout)
    sep_block('base.final_conv.0', 128, 256)
    conv_bn('base.final_conv.1', 'base.final_conv.2', 256, 256, 1)

    for stride, cin in ((8, 64), (16, 128), (32, 256)):
        conv_bn(f'refiner.conv_stride{stride}.0', f'refiner.conv_stride{stride}.1',
                cin, 64, 1, bias=True)
    for stride in (8, 16):
        conv_bn(f'refiner.aggr_stride{stride}.0', f'refiner.aggr_stride{stride}.1',
                64, 64, 3, bias=True)
    for stride in (8, 16, 32):
        p = f'refiner.context_stride{stride}'
        conv_bn(p + '.context_3x3.0', p + '.context_3x3.1', 64, 32, 3, bias=True)
        conv_bn(p + '.dimension_reducer.0', p + '.dimension_reducer.1', 64, 16, 3,
                bias=True)
        conv_bn(p + '.context_5x5.0', p + '.context_5x5.1', 16, 16, 3, bias=True)
        conv_bn(p + '.context_7x7.0', p + '.context_7x7.1', 16, 16, 3, bias=True)
        conv_bn(p + '.context_7x7.3', p + '.context_7x7.4', 16, 16, 3, bias=True)

    for stride in (8, 16, 32):
        w = _conv_w(g, 4, 64, 1, gain=1.0) * cls_gain
        b = torch.zeros(4)
        # channels 2+a are the foreground logits
        b[2:] = cls_bias[(32, 16, 8).index(stride)]
        sd[f'outputs.cls_stride{stride}.weight'] = w
        sd[f'outputs.cls_stride{stride}.bias'] = b
        sd[f'outputs.bbox_stride{stride}.weight'] = _conv_w(g, 8, 64, 1, gain=0.3)
        sd[f'outputs.bbox_stride{stride}.bias'] = torch.randn(8, generator=g) * 0.05
        sd[f'outputs.landmark_stride{stride}.weight'] = _conv_w(g, 20, 64, 1, gain=0.3)
        sd[f'outputs.landmark_stride{stride}.bias'] = torch.randn(20, generator=g) * 0.05
    return sd


# ---------------------------------------------------------------------------
# ArcFace IR-ResNet-100.  Layout: reference arcface/model.py:4-97.
# ---------------------------------------------------------------------------

ARCFACE_UNITS = (3, 13, 30, 3)
ARCFACE_CHANNELS = (64, 64, 128, 256, 512)


def arcface_state_dict(seed=5, units=ARCFACE_UNITS):
    g = _gen(seed)
    sd = {}
    sd['initial_layer.0.weight'] = _conv_w(g, 64, 3, 3)
    _bn(sd, 'initial_layer.1', g, 64)
    sd['initial_layer.2.weight'] = torch.empty(64).uniform_(0.1, 0.4, generator=g)
    for s, n_units in enumerate(units):
        prev_c, curr_c = ARCFACE_CHANNELS[s], ARCFACE_CHANNELS[s + 1]
        for u in range(n_units):
            cin = prev_c if u == 0 else curr_c
            p = f'stages.{s}.{u}'
            _bn(sd, p + '.body.0', g, cin)
            sd[p + '.body.1.weight'] = _conv_w(g, curr_c, cin, 3)
            _bn(sd, p + '.body.2', g, curr_c)
            sd[p + '.body.3.weight'] = torch.empty(curr_c).uniform_(
                0.1, 0.4, generator=g)
            # Residual branch is damped so 49 stacked units keep O(1) range.
            sd[p + '.body.4.weight'] = _conv_w(g, curr_c, curr_c, 3, gain=0.5)
            _bn(sd, p + '.body.5', g, curr_c, gamma=(0.4, 0.6))
            if u == 0:
                sd[p + '.shortcut.0.weight'] = _conv_w(g, curr_c, cin, 1, gain=1.0)
                _bn(sd, p + '.shortcut.1', g, curr_c)
    _bn(sd, 'final_layer.0', g, 512)
    sd['final_layer.3.weight'] = torch.randn(512, 25088, generator=g) / math.sqrt(25088)
    sd['final_layer.3.bias'] = torch.randn(512, generator=g) * 0.05
    _bn(sd, 'final_layer.4', g, 512)
    return sd


# ---------------------------------------------------------------------------
# OpenPose body model.  Layout: reference openpose/model.py:27-112.
# ---------------------------------------------------------------------------

#: (name, cin, cout, k) for the VGG trunk ``model0`` ('P' = 2x2 max-pool).
OPENPOSE_TRUNK = (
    ('conv1_1', 3, 64, 3), ('conv1_2', 64, 64, 3), 'P',
    ('conv2_1', 64, 128, 3), ('conv2_2', 128, 128, 3), 'P',
    ('conv3_1', 128, 256, 3), ('conv3_2', 256, 256, 3), ('conv3_3', 256, 256, 3),
    ('conv3_4', 256, 256, 3), 'P',
    ('conv4_1', 256, 512, 3), ('conv4_2', 512, 512, 3),
    ('conv4_3_CPM', 512, 256, 3), ('conv4_4_CPM', 256, 128, 3),
)


def openpose_stage_layers(stage, branch):
    """[(name, cin, cout, k, relu)] for ``model{stage}_{branch}``.

    ReLU placement follows the reference's ``no_relu_layers`` list including
    its typo: ``Mconv7_stage6_L2`` is NOT in the list, so the final heat-map
    layer keeps its ReLU (openpose/model.py:32-39).
    """
    cout = 38 if branch == 1 else 19
    L = f'L{branch}'
    if stage == 1:
        return [
            (f'conv5_1_CPM_{L}', 128, 128, 3, True),
            (f'conv5_2_CPM_{L}', 128, 128, 3, True),
            (f'conv5_3_CPM_{L}', 128, 128, 3, True),
            (f'conv5_4_CPM_{L}', 128, 512, 1, True),
            (f'conv5_5_CPM_{L}', 512, cout, 1, False),
        ]
    last_relu = (stage == 6 and branch == 2)
    return [
        (f'Mconv1_stage{stage}_{L}', 185, 128, 7, True),
        (f'Mconv2_stage{stage}_{L}', 128, 128, 7, True),
        (f'Mconv3_stage{stage}_{L}', 128, 128, 7, True),
        (f'Mconv4_stage{stage}_{L}', 128, 128, 7, True),
        (f'Mconv5_stage{stage}_{L}', 128, 128, 7, True),
        (f'Mconv6_stage{stage}_{L}', 128, 128, 1, True),
        (f'Mconv7_stage{stage}_{L}', 128, cout, 1, last_relu),
    ]


def openpose_state_dict(seed=7, peaks=False):
    """Seeded OpenPose checkpoint.

    ``peaks=True`` re-calibrates the two output layers (``Mconv7_stage6_L1/L2``) so that the
    decode has real work on uniform-noise frames: every heat-map channel is standardised over a
    fixed calibration batch and mapped to ``relu(0.25 z - 0.35)`` (peaks where a channel is
    ~1.8 sigma above its mean: 10-30 peaks per part and frame), every PAF channel to ``0.5 z``
    (zero mean, so limb directions are random and about a third of the close peak pairs pass
    the line-integral test).  Measured with the oracle on 1080p noise frames: 6-9 humans of
    4-8 joints per frame survive the ``count >= 4, score >= 0.4`` filter.  Without it random
    weights give no peak above 0.1 and peaks/limbs/assembly would be no-ops.
    """
    g = _gen(seed)
    sd = {}

    def conv(prefix, cin, cout, k, gain=math.sqrt(2.0)):
        sd[prefix + '.weight'] = _conv_w(g, cout, cin, k, gain=gain)
        sd[prefix + '.bias'] = torch.randn(cout, generator=g) * 0.05

    for item in OPENPOSE_TRUNK:
        if item == 'P':
            continue
        name, cin, cout, k = item
        conv('model0.' + name, cin, cout, k)
    for stage in range(1, 7):
        for branch in (1, 2):
            layers = openpose_stage_layers(stage, branch)
            for i, (name, cin, cout, k, _relu) in enumerate(layers):
                last = i == len(layers) - 1
                conv(f'model{stage}_{branch}.{name}', cin, cout, k,
                     gain=0.5 if last else math.sqrt(2.0))
    # Final heat-map layer: damped and shifted so that, on noise frames, only a
    # few up-sampled maxima clear the 0.1 peak threshold (tens of peaks on one
    # or two parts, like a real frame) instead of ~150 per part.
    sd['model6_2.Mconv7_stage6_L2.weight'] *= 0.4
    sd['model6_2.Mconv7_stage6_L2.bias'] = sd['model6_2.Mconv7_stage6_L2.bias'] * 0.4 - 0.05
    if peaks:
        _calibrate_openpose_outputs(sd)
    return sd


def _openpose_penultimate(sd, x):
    """Inputs of the two output layers (``Mconv7_stage6_L{1,2}``) for a batch ``x`` — a plain
    torch CPU fp32 evaluation of the checkpoint, used once to calibrate synthetic weights."""
    import torch.nn.functional as F
    with torch.no_grad():
        out = x
        for item in OPENPOSE_TRUNK:
            if item == 'P':
                out = F.max_pool2d(out, 2, 2, 0)
                continue
            name, _cin, _cout, k = item
            out = F.relu(F.conv2d(out, sd[f'model0.{name}.weight'], sd[f'model0.{name}.bias'],
                                  padding=k // 2))
        trunk = inp = out
        for stage in range(1, 7):
            outs, pen = [], []
            for branch in (1, 2):
                y = inp
                layers = openpose_stage_layers(stage, branch)
                for li, (name, _cin, _cout, k, relu) in enumerate(layers):
                    if li == len(layers) - 1:
                        pen.append(y)
                    pfx = f'model{stage}_{branch}.{name}'
                    y = F.conv2d(y, sd[pfx + '.weight'], sd[pfx + '.bias'], padding=k // 2)
                    if relu:
                        y = F.relu(y)
                outs.append(y)
            inp = torch.cat([outs[0], outs[1], trunk], dim=1)
    return pen


def _calibrate_openpose_outputs(sd, paf_gain=0.5, heat_gain=0.25, heat_bias=-0.35):
    import torch.nn.functional as F
    g = _gen(4242)
    big = torch.randint(0, 256, (2, 3, 540, 960), generator=g).float()
    # short side 184 like the pose wrapper; bilinear like cv2 (statistics only, not bit-exact)
    x = F.interpolate(big, size=(184, 327), mode='bilinear', align_corners=False).round() / 255.0 - 0.5
    pen = _openpose_penultimate(sd, x)
    for branch, (gain, bias) in ((1, (paf_gain, 0.0)), (2, (heat_gain, heat_bias))):
        pfx = f'model6_{branch}.Mconv7_stage6_L{branch}'
        raw = F.conv2d(pen[branch - 1], sd[pfx + '.weight'])
        mean, std = raw.mean((0, 2, 3)), raw.std((0, 2, 3))
        sd[pfx + '.weight'] = sd[pfx + '.weight'] * (gain / std).view(-1, 1, 1, 1)
        sd[pfx + '.bias'] = -mean * gain / std + bias
