"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of the
three reference ``nn.Module`` graphs, written functionally over a reference
``state_dict``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this.  Pinned against the reference
modules themselves by ``oracle/make_golden.py`` (run in the build container,
where ``/root/reference`` is importable) and against the committed fixtures in
``tests/golden/`` everywhere else.

Follows (relative to the reference tree):
  * RetinaFace  ``terran/face/detection/retinaface/model.py:53-341``
  * ArcFace     ``terran/face/recognition/arcface/model.py:4-97``
  * OpenPose    ``terran/pose/openpose/model.py:27-141``
"""
import torch
import torch.nn.functional as F

from terran_b200.synth import (
    ARCFACE_UNITS, OPENPOSE_TRUNK, RETINAFACE_SCALES, openpose_stage_layers,
)


def _bn(sd, p, x, eps):
    return F.batch_norm(
        x, sd[p + '.running_mean'], sd[p + '.running_var'],
        sd[p + '.weight'], sd[p + '.bias'], training=False, eps=eps,
    )


# ------------------------------------------------------------------ RetinaFace

def _cbr(sd, pc, pb, x, eps, stride=1, padding=0, groups=1):
    """conv -> BN -> ReLU (model.py:26-39, 126-153, 178-202)."""
    x = F.conv2d(x, sd[pc + '.weight'], sd.get(pc + '.bias'), stride=stride,
                 padding=padding, groups=groups)
    return F.relu(_bn(sd, pb, x, eps))


def _sep_block(sd, p, x, stride):
    conv = _cbr(sd, p + '.conv_block.0', p + '.conv_block.1', x, 1e-5)
    sep = _cbr(sd, p + '.sep_block.0', p + '.sep_block.1', conv, 1e-5,
               stride=stride, padding=1, groups=conv.shape[1])
    return conv, sep


def _context(sd, p, x):
    """ContextModule (model.py:155-165): cat[3x3(32), 5x5(16), 7x7(16)]."""
    e = 2e-5
    red = _cbr(sd, p + '.dimension_reducer.0', p + '.dimension_reducer.1', x, e, padding=1)
    c3 = _cbr(sd, p + '.context_3x3.0', p + '.context_3x3.1', x, e, padding=1)
    c5 = _cbr(sd, p + '.context_5x5.0', p + '.context_5x5.1', red, e, padding=1)
    c7 = _cbr(sd, p + '.context_7x7.0', p + '.context_7x7.1', red, e, padding=1)
    c7 = _cbr(sd, p + '.context_7x7.3', p + '.context_7x7.4', c7, e, padding=1)
    return torch.cat([c3, c5, c7], dim=1)


def retinaface_features(sd, x):
    """Backbone + refiner: returns the three 64-channel context maps
    (stride 8, 16, 32).  ``x`` is (N,3,H,W) fp32 **BGR** raw 0..255."""
    x = x.contiguous()
    out = _cbr(sd, 'base.first_conv_block.0', 'base.first_conv_block.1', x, 1e-5,
               stride=2, padding=1)
    out = _cbr(sd, 'base.first_conv_block.3', 'base.first_conv_block.4', out, 1e-5,
               padding=1, groups=8)
    taps = []
    for si, blocks in enumerate(RETINAFACE_SCALES):
        for bi, (_cin, _cout, stride) in enumerate(blocks):
            conv, out = _sep_block(sd, f'base.scales.{si}.{bi}', out, stride)
        taps.append(conv)      # the 1x1 output of the last block (model.py:45-46)
    _, out = _sep_block(sd, 'base.final_conv.0', out, 1)
    out = _cbr(sd, 'base.final_conv.1', 'base.final_conv.2', out, 1e-5)
    taps.append(out)

    e = 2e-5
    p8 = _cbr(sd, 'refiner.conv_stride8.0', 'refiner.conv_stride8.1', taps[0], e)
    p16 = _cbr(sd, 'refiner.conv_stride16.0', 'refiner.conv_stride16.1', taps[1], e)
    p32 = _cbr(sd, 'refiner.conv_stride32.0', 'refiner.conv_stride32.1', taps[2], e)
    up = F.interpolate(p32, scale_factor=2)[:, :, :p16.shape[2], :p16.shape[3]]
    p16 = _cbr(sd, 'refiner.aggr_stride16.0', 'refiner.aggr_stride16.1', p16 + up, e,
               padding=1)
    up = F.interpolate(p16, scale_factor=2)[:, :, :p8.shape[2], :p8.shape[3]]
    p8 = _cbr(sd, 'refiner.aggr_stride8.0', 'refiner.aggr_stride8.1', p8 + up, e,
              padding=1)
    return [
        _context(sd, 'refiner.context_stride8', p8),
        _context(sd, 'refiner.context_stride16', p16),
        _context(sd, 'refiner.context_stride32', p32),   # un-aggregated (model.py:243)
    ]


def retinaface_heads(sd, ctx):
    """OutputsPredictor (model.py:258-316): 9 tensors, order s32, s16, s8;
    class scores soft-maxed over the channel pairs (a, a+2)."""
    out = []
    for stride, f in ((32, ctx[2]), (16, ctx[1]), (8, ctx[0])):
        cls = F.conv2d(f, sd[f'outputs.cls_stride{stride}.weight'],
                       sd[f'outputs.cls_stride{stride}.bias'])
        n, a, h, w = cls.shape
        prob = F.softmax(cls.reshape(n, 2, -1, w), dim=1).reshape(n, a, h, w)
        bbox = F.conv2d(f, sd[f'outputs.bbox_stride{stride}.weight'],
                        sd[f'outputs.bbox_stride{stride}.bias'])
        lmk = F.conv2d(f, sd[f'outputs.landmark_stride{stride}.weight'],
                       sd[f'outputs.landmark_stride{stride}.bias'])
        out += [prob, bbox, lmk]
    return out


def retinaface_forward(sd, x):
    with torch.no_grad():
        return retinaface_heads(sd, retinaface_features(sd, x))


# --------------------------------------------------------------------- ArcFace

def arcface_forward(sd, x, units=ARCFACE_UNITS):
    """FaceResNet100.forward (arcface/model.py:87-97).  ``x`` is (N,3,112,112)
    fp32 BGR raw 0..255; returns the (N,512) un-normalised embedding."""
    e = 2e-5
    with torch.no_grad():
        out = (x - 127.5) * 0.0078125
        out = F.conv2d(out, sd['initial_layer.0.weight'], padding=1)
        out = F.prelu(_bn(sd, 'initial_layer.1', out, e), sd['initial_layer.2.weight'])
        for s, n_units in enumerate(units):
            for u in range(n_units):
                p = f'stages.{s}.{u}'
                stride = 2 if u == 0 else 1
                b = _bn(sd, p + '.body.0', out, e)
                b = F.conv2d(b, sd[p + '.body.1.weight'], padding=1)
                b = F.prelu(_bn(sd, p + '.body.2', b, e), sd[p + '.body.3.weight'])
                b = F.conv2d(b, sd[p + '.body.4.weight'], stride=stride, padding=1)
                b = _bn(sd, p + '.body.5', b, e)
                if u == 0:
                    sc = F.conv2d(out, sd[p + '.shortcut.0.weight'], stride=stride)
                    sc = _bn(sd, p + '.shortcut.1', sc, e)
                else:
                    sc = out
                out = b + sc
        out = _bn(sd, 'final_layer.0', out, e)
        out = out.flatten(1)                       # (C,H,W) order
        out = F.linear(out, sd['final_layer.3.weight'], sd['final_layer.3.bias'])
        out = F.batch_norm(out, sd['final_layer.4.running_mean'],
                           sd['final_layer.4.running_var'], sd['final_layer.4.weight'],
                           sd['final_layer.4.bias'], training=False, eps=e)
    return out


# -------------------------------------------------------------------- OpenPose

def openpose_forward(sd, x):
    """BodyPoseModel.forward (openpose/model.py:114-141).  ``x`` is (N,3,H,W)
    fp32 RGB in [-0.5, 0.5]; returns (PAF (N,38,h,w), heat (N,19,h,w))."""
    with torch.no_grad():
        out = x
        for item in OPENPOSE_TRUNK:
            if item == 'P':
                out = F.max_pool2d(out, 2, 2, 0)
                continue
            name, _cin, _cout, k = item
            out = F.relu(F.conv2d(out, sd[f'model0.{name}.weight'],
                                  sd[f'model0.{name}.bias'], padding=k // 2))
        trunk = out
        inp = trunk
        branches = None
        for stage in range(1, 7):
            branches = []
            for branch in (1, 2):
                y = inp
                for name, _cin, _cout, k, relu in openpose_stage_layers(stage, branch):
                    pfx = f'model{stage}_{branch}.{name}'
                    y = F.conv2d(y, sd[pfx + '.weight'], sd[pfx + '.bias'],
                                 padding=k // 2)
                    if relu:
                        y = F.relu(y)
                branches.append(y)
            inp = torch.cat([branches[0], branches[1], trunk], dim=1)
    return branches[0], branches[1]
