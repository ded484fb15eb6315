"""TEST INFRASTRUCTURE: the layer-program builders restated in Python (numpy/torch), kept as an
independent check of the library's own builder (``terran_b200/csrc/program.cu``):
``tests/test_programs_cpu.py`` asserts that both produce the same op list and — bit for bit for
the fp16 filters, to float round-off for the folded fp32 vectors — the same weight blob.  Not
imported by the product.
"""
import numpy as np
import torch

from terran_b200 import _native as nat
from terran_b200.synth import ARCFACE_CHANNELS, ARCFACE_UNITS, OPENPOSE_TRUNK, RETINAFACE_SCALES, \
    openpose_stage_layers
from terran_b200.weights import Program


def _r(x, m):
    return (x + m - 1) // m * m


def bn_fold(sd, prefix, eps, conv_bias=None):
    """(scale, shift) of BN(conv + bias) as an affine of the conv output."""
    g, b = sd[prefix + '.weight'].double(), sd[prefix + '.bias'].double()
    m, v = sd[prefix + '.running_mean'].double(), sd[prefix + '.running_var'].double()
    scale = g / torch.sqrt(v + eps)
    bias = conv_bias.double() if conv_bias is not None else 0.0
    shift = b + (bias - m) * scale
    return scale.float().numpy(), shift.float().numpy()


def pre_bn_fold(w, s0, t0, s1, t1):
    """Fold a per-channel affine ``x*s0 + t0`` that PRECEDES a zero-padded 3x3 stride-1 conv
    (and the affine ``y*s1 + t1`` that follows it) into the conv:

        conv_W(pad(x*s0 + t0)) * s1 + t1  =  conv_{W*s0}(pad(x)) * s1 + shift9[class]

    Padding is applied AFTER the first affine, so its shift only reaches the taps that are in
    bounds: the input-independent term ``sum_{taps in bounds} W[:, :, tap] @ t0`` depends on the
    output pixel's border class (3 row classes x 3 column classes).  Returns (w * s0,
    shift9 (9, cout)) with class = 3*row_class + col_class, 0 = first, 1 = inner, 2 = last."""
    w = w.detach().double()
    s0, t0 = torch.as_tensor(s0).double(), torch.as_tensor(t0).double()
    s1, t1 = torch.as_tensor(s1).double(), torch.as_tensor(t1).double()
    per_tap = torch.einsum('oikl,i->okl', w, t0)            # (cout, 3, 3)
    shift9 = torch.empty((9, w.shape[0]), dtype=torch.float64)
    rows = {0: (1, 2), 1: (0, 1, 2), 2: (0, 1)}             # filter rows that stay in bounds
    for rc, rs in rows.items():
        for cc, cs in rows.items():
            term = per_tap[:, list(rs)][:, :, list(cs)].sum((1, 2))
            shift9[rc * 3 + cc] = term * s1 + t1
    return (w * s0.view(1, -1, 1, 1)).float(), shift9.float().numpy()


# ------------------------------------------------------------------ RetinaFace

def retinaface_program(sd, fused=True):
    """Program + roles for reference ``RetinaFace`` (retinaface/model.py:319-341).
    The stem reads the frame in MODEL channel order (BGR); callers with RGB
    memory pass a pointer to channel 2 and a channel stride of -1.

    ``fused`` (default: env ``TRB_RETINA_FUSED`` != 0): every depthwise 3x3 of the backbone
    runs inside the 1x1 conv that consumes it (one ``TR_OP_SEPCONV`` per pair) and the
    refiner / context / head convs go to the warp-level mma.sync kernel; otherwise one op
    per reference layer on the tcgen05 / direct kernels."""
    P = Program()
    relu = nat.TR_ACT_RELU
    engine = nat.TR_ENGINE_MMA if fused else nat.TR_ENGINE_AUTO

    def cbr(pc, pb, in_, out, eps, **kw):
        s, t = bn_fold(sd, pb, eps, sd.get(pc + '.bias'))
        w = sd[pc + '.weight']
        # measured (profiles/r01_retinaface_mma.txt): the warp-level kernel wins where either
        # channel count is <= 16; the 64-channel 3x3 / lateral 1x1 layers stay on tcgen05
        eng = engine if min(w.shape[0], w.shape[1]) <= 16 else nat.TR_ENGINE_AUTO
        P.conv(w, s, t, in_, out, act=relu, engine=eng, **kw)

    b = P.buffer(8)
    s, t = bn_fold(sd, 'base.first_conv_block.1', 1e-5)
    P.stem(sd['base.first_conv_block.0.weight'], s, t, b, stride=2, act=relu)

    # The backbone is stem -> dw -> [1x1 -> dw]* -> 1x1 (ConvSepBlock = 1x1 conv_block then
    # depthwise sep_block, model.py:6-50).  Pair every depthwise with the 1x1 that FOLLOWS it.
    blocks = [(f'base.scales.{si}.{bi}', cout, stride)
              for si, bl in enumerate(RETINAFACE_SCALES) for bi, (_cin, cout, stride) in enumerate(bl)]
    blocks.append(('base.final_conv.0', 256, 1))
    tap_after = {len(RETINAFACE_SCALES[0]) - 1, len(RETINAFACE_SCALES[0]) + len(RETINAFACE_SCALES[1]) - 1}
    pending = ('base.first_conv_block.3', 'base.first_conv_block.4', 1)   # (dw conv, dw bn, stride)
    x, ch = b, 8
    taps = []
    pointwise = [(p + '.conv_block.0', p + '.conv_block.1', c) for p, c, _ in blocks]
    pointwise.append(('base.final_conv.1', 'base.final_conv.2', 256))
    next_dw = [(p + '.sep_block.0', p + '.sep_block.1', st) for p, _, st in blocks] + [None]
    for i, ((pc, pb, cout), nxt) in enumerate(zip(pointwise, next_dw)):
        dwc, dwb, stride = pending
        ds, dt = bn_fold(sd, dwb, 1e-5)
        s, t = bn_fold(sd, pb, 1e-5)
        y = P.buffer(cout)
        if fused:
            P.sepconv(sd[dwc + '.weight'], ds, dt, sd[pc + '.weight'], s, t, x, y, stride=stride)
        else:
            d = P.buffer(ch)
            P.dwconv(sd[dwc + '.weight'], ds, dt, x, d, stride=stride)
            P.conv(sd[pc + '.weight'], s, t, d, y, act=relu)
        if i in tap_after:
            taps.append(y)
        x, ch, pending = y, cout, nxt
    c32 = x
    c8, c16 = taps

    e = 2e-5
    p32 = P.buffer(64)
    cbr('refiner.conv_stride32.0', 'refiner.conv_stride32.1', c32, p32, e)
    p16 = P.buffer(64)
    cbr('refiner.conv_stride16.0', 'refiner.conv_stride16.1', c16, p16, e, res=p32, res_up2=1)
    a16 = P.buffer(64)
    cbr('refiner.aggr_stride16.0', 'refiner.aggr_stride16.1', p16, a16, e)
    p8 = P.buffer(64)
    cbr('refiner.conv_stride8.0', 'refiner.conv_stride8.1', c8, p8, e, res=a16, res_up2=1)
    a8 = P.buffer(64)
    cbr('refiner.aggr_stride8.0', 'refiner.aggr_stride8.1', p8, a8, e, sync=nat.TR_SYNC_FORK)

    heads, ctxs = {}, {}
    for stride, feat in ((8, a8), (16, a16), (32, p32)):
        p = f'refiner.context_stride{stride}'
        ctx = P.buffer(64)
        red = P.buffer(16)
        tmp = P.buffer(16)
        lane = 0 if stride == 8 else 1        # strides 16 and 32 on the side stream
        cbr(p + '.context_3x3.0', p + '.context_3x3.1', feat, ctx, e, out_coff=0, lane=lane)
        cbr(p + '.dimension_reducer.0', p + '.dimension_reducer.1', feat, red, e, lane=lane)
        cbr(p + '.context_5x5.0', p + '.context_5x5.1', red, ctx, e, out_coff=32, lane=lane)
        cbr(p + '.context_7x7.0', p + '.context_7x7.1', red, tmp, e, lane=lane)
        cbr(p + '.context_7x7.3', p + '.context_7x7.4', tmp, ctx, e, out_coff=48, lane=lane)
        # fused head: [4 class logits | 8 bbox | 20 landmark] -> fp32
        w = torch.cat([sd[f'outputs.cls_stride{stride}.weight'],
                       sd[f'outputs.bbox_stride{stride}.weight'],
                       sd[f'outputs.landmark_stride{stride}.weight']], 0)
        bias = torch.cat([sd[f'outputs.cls_stride{stride}.bias'],
                          sd[f'outputs.bbox_stride{stride}.bias'],
                          sd[f'outputs.landmark_stride{stride}.bias']], 0)
        head = P.buffer(32, f32=True)
        P.conv(w, np.ones(32, np.float32), bias.float().numpy(), ctx, head, engine=engine, lane=lane)
        heads[stride] = head
        ctxs[stride] = ctx
    roles = {'heads': [heads[32], heads[16], heads[8]], 'context': ctxs}
    return P, roles


# --------------------------------------------------------------------- ArcFace

def arcface_program(sd, units=ARCFACE_UNITS):
    """Program for reference ``FaceResNet100`` (arcface/model.py:38-97).  Input
    in model channel order (BGR).

    Every ``Unit`` starts with a BatchNorm in front of a zero-padded 3x3 conv
    (``body[0]``, ``body[1]``; model.py:11-14).  It is folded into that conv exactly — scale
    into the filters, shift into nine border-class shift vectors (``pre_bn_fold``) — so the
    residual stream ``x`` is the only tensor a unit reads and writes: no second, normalised
    copy of every activation."""
    P = Program()
    e = 2e-5

    C0 = ARCFACE_CHANNELS[0]
    x = P.buffer(C0)
    s, t = bn_fold(sd, 'initial_layer.1', e)
    P.stem(sd['initial_layer.0.weight'], s, t, x, stride=1, act=nat.TR_ACT_PRELU,
           slope=sd['initial_layer.2.weight'].float().numpy(),
           in_scale=0.0078125, in_shift=-127.5 * 0.0078125)

    n_stages = len(units)
    for si, n_units in enumerate(units):
        cout = ARCFACE_CHANNELS[si + 1]
        y_full = P.buffer(cout)      # conv1 output of the strided first unit (full res)
        y = P.buffer(cout)
        sc = P.buffer(cout)
        xs = [P.buffer(cout), P.buffer(cout)]
        for u in range(n_units):
            p = f'stages.{si}.{u}'
            stride = 2 if u == 0 else 1
            s0, t0 = bn_fold(sd, p + '.body.0', e)
            s1, t1 = bn_fold(sd, p + '.body.2', e)
            w1, shift9 = pre_bn_fold(sd[p + '.body.1.weight'], s0, t0, s1, t1)
            yb = y_full if u == 0 else y
            P.conv(w1, s1, shift9[4], x, yb, act=nat.TR_ACT_PRELU,
                   slope=sd[p + '.body.3.weight'].float().numpy(), shift9=shift9)
            if u == 0:
                s, t = bn_fold(sd, p + '.shortcut.1', e)
                P.conv(sd[p + '.shortcut.0.weight'], s, t, x, sc, stride=2)
                res = sc
            else:
                res = x
            x_new = xs[u & 1]
            s, t = bn_fold(sd, p + '.body.5', e)
            P.conv(sd[p + '.body.4.weight'], s, t, yb, x_new, stride=stride, res=res)
            x = x_new

    # final_layer: BN2d (no padding follows -> folded into the FC exactly),
    # Flatten in (C,H,W) order -> permuted to our (H,W,C), Linear, BN1d.
    Cl = ARCFACE_CHANNELS[len(units)]
    s0, t0 = (v.astype(np.float64) for v in bn_fold(sd, 'final_layer.0', e))
    W = sd['final_layer.3.weight'].double().numpy()
    hw = W.shape[1] // Cl
    side = int(round(hw ** 0.5))
    W = W.reshape(512, Cl, side, side)
    bias = sd['final_layer.3.bias'].double().numpy() + (W * t0[None, :, None, None]).sum((1, 2, 3))
    Wf = (W * s0[None, :, None, None]).transpose(0, 2, 3, 1).reshape(512, hw * Cl, 1, 1)
    g = sd['final_layer.4.weight'].double().numpy()
    b = sd['final_layer.4.bias'].double().numpy()
    m = sd['final_layer.4.running_mean'].double().numpy()
    v = sd['final_layer.4.running_var'].double().numpy()
    scale = g / np.sqrt(v + e)
    shift = b + (bias - m) * scale
    flat = P.buffer(hw * Cl)
    P.view(x, flat)
    emb = P.buffer(512, f32=True)
    P.conv(torch.from_numpy(Wf.astype(np.float32)), scale.astype(np.float32),
           shift.astype(np.float32), flat, emb, k=1, pad=0)
    return P, {'embedding': emb}


# -------------------------------------------------------------------- OpenPose

#: position of reference concat channel c (cat[PAF 38, heat 19, trunk 128]) in
#: the padded 192-channel buffer [PAF 0..37 | pad | heat 40..58 | pad | trunk 64..191]
OPENPOSE_CAT_MAP = np.concatenate([np.arange(38), 40 + np.arange(19), 64 + np.arange(128)])


def openpose_program(sd):
    """Program for reference ``BodyPoseModel`` (openpose/model.py:27-141).  The
    frame is read as stored (RGB, no flip: wrapper.py:116-122)."""
    P = Program()
    relu = nat.TR_ACT_RELU

    def one(cout):
        return np.ones(cout, np.float32)

    cat = [P.buffer(192), P.buffer(192)]
    x = None
    ch = 3
    items = list(OPENPOSE_TRUNK)
    for i, item in enumerate(items):
        if item == 'P':
            y = P.buffer(ch)
            P.maxpool(x, y, ch)
            x = y
            continue
        name, cin, cout, _k = item
        w, b = sd[f'model0.{name}.weight'], sd[f'model0.{name}.bias'].float().numpy()
        if cin == 3:
            y = P.buffer(cout)
            # reference: x.astype(f32) / 255.0 - 0.5 (wrapper.py:116-122)
            P.stem(w, one(cout), b, y, stride=1, act=relu, in_scale=1.0 / 255.0, in_shift=-0.5)
        elif i == len(items) - 1:
            P.conv(w, one(cout), b, x, cat[0], out_coff=64, act=relu)
            P.copy(cat[0], 64, cat[1], 64, 128)
            y = None
        else:
            y = P.buffer(cout)
            P.conv(w, one(cout), b, x, y, act=relu)
        x, ch = y, cout

    for stage in range(1, 7):
        src = cat[0] if stage == 1 else cat[stage % 2]
        dst = cat[0] if stage == 1 else cat[(stage + 1) % 2]
        specs = {b: openpose_stage_layers(stage, b) for b in (1, 2)}
        n_layers = len(specs[1])
        tmp = [P.buffer(256), P.buffer(256)]
        xb, ch = None, 0
        # layer 0: one conv with both filter banks stacked (same input); layers 1..n-2: ONE grouped
        # conv (groups = 2) over the [branch 1 | branch 2] channel pair; last layer: per branch
        for li in range(n_layers - 1):
            (n1, cin, c1, k, _), (n2, _, c2, _, _) = specs[1][li], specs[2][li]
            w = torch.cat([sd[f'model{stage}_1.{n1}.weight'], sd[f'model{stage}_2.{n2}.weight']], 0)
            b = torch.cat([sd[f'model{stage}_1.{n1}.bias'], sd[f'model{stage}_2.{n2}.bias']], 0)
            kw = {}
            sync = 0
            if li == 0:
                kw = dict(in_coff=64) if stage == 1 else dict(in_map=OPENPOSE_CAT_MAP, cin_pad=192)
            else:
                kw = dict(groups=2)
            y = P.buffer(c1 + c2) if c1 + c2 != 256 else tmp[li & 1]
            P.conv(w, one(c1 + c2), b.float().numpy(), src if li == 0 else xb, y, act=relu, sync=sync, **kw)
            xb, ch = y, c1
        # last layers: one 1x1 conv over the pair, output rows = concat channels 0..63
        # (rows 0..37 PAF filters on the first ch inputs, rows 40..58 heat filters on the second)
        (n1, _, c1, _, relu1), (n2, _, c2, _, relu2) = specs[1][-1], specs[2][-1]
        w1, w2 = sd[f'model{stage}_1.{n1}.weight'].float(), sd[f'model{stage}_2.{n2}.weight'].float()
        wm = torch.zeros((64, 2 * ch, 1, 1))
        wm[:c1, :ch] = w1
        wm[40:40 + c2, ch:] = w2
        bm = np.zeros(64, np.float32)
        bm[:c1] = sd[f'model{stage}_1.{n1}.bias'].float().numpy()
        bm[40:40 + c2] = sd[f'model{stage}_2.{n2}.bias'].float().numpy()
        kw = {}
        if relu1 or relu2:
            slope = np.ones(64, np.float32)
            if relu1:
                slope[:c1] = 0
            if relu2:
                slope[40:40 + c2] = 0
            kw = dict(act=nat.TR_ACT_PRELU, slope=slope)
        P.conv(wm, one(64), bm, xb, dst, **kw)
        P.ops[-1].cin_real, P.ops[-1].cout_real = ch, c1 + c2
    return P, {'maps': cat[(6 + 1) % 2], 'paf_coff': 0, 'heat_coff': 40}


