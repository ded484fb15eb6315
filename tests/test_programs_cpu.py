"""CPU: the layer programs emitted by ``terran_b200/weights.py`` — executed by the torch
interpreter in ``tests/program_interpreter.py`` exactly as the C ABI specifies the ops — agree
with the oracle's restatement of the reference ``nn.Module`` graphs.  This pins the whole
weight-ingestion path (checkpoint key layout, BN folding, filter re-indexing, fused heads,
depthwise+1x1 pairing, concat-free channel slices) without a GPU."""
import numpy as np
import pytest
import torch

from oracle import nets
from terran_b200 import synth, weights
from tests.program_interpreter import Interpreter


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'unfused'])
def test_retinaface_program_matches_oracle(fused):
    """retinaface/model.py:53-341 vs the 13 TR_OP_SEPCONV (or 13 + 13 unfused) backbone ops,
    the FPN with up-sampled residuals, the context slices and the fused fp32 heads."""
    sd = synth.retinaface_state_dict()
    P, roles = weights.retinaface_program(sd, fused=fused)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (2, 64, 96, 3), dtype=np.uint8))     # model order (BGR)
    bufs = Interpreter(P, round_activations=False).run(img)
    want = nets.retinaface_forward(sd, img.float().permute(0, 3, 1, 2))
    for li, head in enumerate(roles['heads']):                                      # s32, s16, s8
        got = bufs[head].permute(0, 3, 1, 2)                                        # (N, 32, h, w)
        prob, bbox, lmk = want[3 * li:3 * li + 3]
        assert got.shape[2:] == bbox.shape[2:]
        # fp16-rounded weights, fp32 activations: a few 1e-3 of the tensor's scale
        assert _rel(got[:, 4:12], bbox) < 5e-3
        assert _rel(got[:, 12:32], lmk) < 5e-3
        n, a, h, w = got[:, :4].shape
        p = torch.softmax(got[:, :4].reshape(n, 2, -1, w), dim=1).reshape(n, a, h, w)
        assert (p - prob).abs().max() < 2e-2                                        # class gain x30 in the synthetic heads


def test_retinaface_fused_and_unfused_programs_agree():
    sd = synth.retinaface_state_dict()
    rng = np.random.default_rng(1)
    img = torch.from_numpy(rng.integers(0, 256, (1, 70, 90, 3), dtype=np.uint8))
    outs = []
    for fused in (True, False):
        P, roles = weights.retinaface_program(sd, fused=fused)
        bufs = Interpreter(P).run(img)                 # fp16 activations like the kernels
        outs.append([bufs[h] for h in roles['heads']])
    for a, b in zip(*outs):
        # same arithmetic except the depthwise filters (fp16 copy in the fused kernel)
        assert _rel(a[..., 4:], b[..., 4:]) < 5e-3


def test_openpose_program_matches_oracle():
    """openpose/model.py:27-141: merged first layers of the two branches, concat written in
    place into the 192-channel buffer [PAF | pad | heat | pad | trunk], re-indexed 7x7 filters."""
    sd = synth.openpose_state_dict()
    P, roles = weights.openpose_program(sd)
    rng = np.random.default_rng(2)
    img = torch.from_numpy(rng.integers(0, 256, (1, 48, 64, 3), dtype=np.uint8))     # RGB as stored
    bufs = Interpreter(P, round_activations=False).run(img)
    x = img.float().permute(0, 3, 1, 2) / 255.0 - 0.5
    paf, heat = nets.openpose_forward(sd, x)
    maps = bufs[roles['maps']].permute(0, 3, 1, 2)
    assert _rel(maps[:, roles['paf_coff']:roles['paf_coff'] + 38], paf) < 5e-3
    assert _rel(maps[:, roles['heat_coff']:roles['heat_coff'] + 19], heat) < 5e-3


def test_arcface_program_matches_oracle():
    """arcface/model.py:4-97 at reduced depth: input affine on in-bounds taps, pre-conv BN
    carried as the producer's second output, shortcut convs, BN2d folded into the permuted FC."""
    units = (1, 2, 1, 1)
    sd = synth.arcface_state_dict(units=units)
    P, roles = weights.arcface_program(sd, units=units)
    rng = np.random.default_rng(3)
    img = torch.from_numpy(rng.integers(0, 256, (2, 112, 112, 3), dtype=np.uint8))   # model order (BGR)
    bufs = Interpreter(P, round_activations=False).run(img)
    want = nets.arcface_forward(sd, img.float().permute(0, 3, 1, 2), units=units)
    got = bufs[roles['embedding']].reshape(2, 512)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    assert (cos > 0.9999).all(), cos
    assert _rel(got, want) < 5e-3


def test_program_traffic_model_of_the_retinaface_stack():
    """Algorithmic bytes per frame at the benchmark shape (32 x 416 x 739): the fused program
    must be close to SURVEY.md 8d's block-fused figure (20.1 MB/frame) and well below the
    layer-wise one (42.6 MB/frame), which the unfused program reproduces."""
    sd = synth.retinaface_state_dict()
    fused, _ = weights.retinaface_program(sd, fused=True)
    plain, _ = weights.retinaface_program(sd, fused=False)
    bf, per_op = weights.program_traffic(fused, 32, 416, 739)
    bp, _ = weights.program_traffic(plain, 32, 416, 739)
    mb_f, mb_p = bf / 32 / 1e6, bp / 32 / 1e6
    assert 18.0 < mb_f < 30.0, mb_f
    assert 38.0 < mb_p < 48.0, mb_p
    assert len(per_op) == len(fused.ops) and all(r > 0 and w > 0 for _, r, w in per_op)
    # shapes inferred like the executor: three head maps of ceil(H/stride) x ceil(W/stride)
    _, per = weights.program_traffic(fused, 1, 416, 739)
    heads = [w for (i, _, w) in per if fused.buffers[fused.ops[i].out][1]]
    assert sorted(heads) == [13 * 24 * 32 * 4, 26 * 47 * 32 * 4, 52 * 93 * 32 * 4]


# ---- the library's builder (csrc/program.cu) against its Python restatement

def _compare_programs(P, Q):
    """Same buffers, same ops (every field incl. blob offsets), fp16 filter blocks bit-equal,
    fp32 vectors equal to float round-off (C++ and numpy sum in different orders)."""
    assert P.buffers == Q.buffers
    assert len(P.ops) == len(Q.ops)
    names = [f[0] for f in type(P.ops[0])._fields_]
    for i, (a, b) in enumerate(zip(P.ops, Q.ops)):
        for n in names:
            va, vb = getattr(a, n), getattr(b, n)
            if isinstance(va, float):
                assert abs(va - vb) <= 1e-7 * max(1.0, abs(vb)), (i, n, va, vb)
            else:
                assert va == vb, (i, n, va, vb)
    assert len(P.blob) == len(Q.blob)
    pa, qa = np.frombuffer(bytes(P.blob), np.uint8), np.frombuffer(bytes(Q.blob), np.uint8)
    same = pa == qa
    if same.all():
        return 1.0
    # differing bytes may only sit in fp32 regions and differ by round-off
    n4 = len(pa) // 4 * 4
    fa, fb = pa[:n4].view(np.float32), qa[:n4].view(np.float32)
    bad = np.flatnonzero(fa.view(np.uint32) != fb.view(np.uint32))
    assert np.allclose(fa[bad], fb[bad], rtol=2e-6, atol=1e-7), np.abs(fa[bad] - fb[bad]).max()
    return float(same.mean())


@pytest.mark.parametrize('model', ['retinaface', 'retinaface-unfused', 'arcface', 'openpose'])
def test_native_program_builder_matches_python_restatement(native, model):
    from tests import reference_programs as ref
    if model.startswith('retinaface'):
        sd = synth.retinaface_state_dict()
        fused = model == 'retinaface'
        P, roles = weights.retinaface_program(sd, fused=fused)
        Q, qroles = ref.retinaface_program(sd, fused=fused)
        assert roles['heads'] == qroles['heads']
    elif model == 'arcface':
        units = (1, 2, 1, 1)
        sd = synth.arcface_state_dict(units=units)
        P, roles = weights.arcface_program(sd)
        Q, qroles = ref.arcface_program(sd, units=units)
        assert roles['embedding'] == qroles['embedding']
    else:
        sd = synth.openpose_state_dict()
        P, roles = weights.openpose_program(sd)
        Q, qroles = ref.openpose_program(sd)
        assert roles == {k: qroles[k] for k in roles}
    frac = _compare_programs(P, Q)
    assert frac > 0.999


def test_state_dict_blob_errors(native):
    import ctypes as C
    from terran_b200 import _native as nat
    h = C.c_void_p()
    bad = np.frombuffer(b'NOPE' + b'\0' * 12, np.uint8)
    assert nat.lib().tr_program_build(b'retinaface', C.c_void_p(bad.ctypes.data), len(bad), 0, C.byref(h)) != 0
    assert b'magic' in nat.lib().tr_last_error()
    with pytest.raises(nat.NativeError, match="no tensor 'base.first_conv_block.1.weight'"):
        weights.native_program('retinaface', {'x': torch.zeros(3)})
    with pytest.raises(nat.NativeError, match='unknown model'):
        weights.native_program('resnet', {'x': torch.zeros(3)})
