"""ORACLE (test infrastructure, not product code): numpy restatement of the
RetinaFace post-processing — anchors, bbox / landmark decode, score threshold,
descending sort and greedy NMS — plus the ``Detection`` wrapper's host-side
rescale/round.

Follows (relative to the reference tree):
  * anchors            ``terran/face/detection/retinaface/anchors.py:7-134``
  * slicing + decode   ``terran/face/detection/retinaface/wrapper.py:25-89,169-202``
  * threshold/sort/NMS ``terran/face/detection/retinaface/wrapper.py:207-236``
    (``torchvision.ops.nms`` CPU semantics: IoU in f32, compared as a double
    against the threshold, areas without the +1)
  * rescale + round    ``terran/face/detection/__init__.py:59-84``

Pinned by ``tests/golden/retinaface_*.npz`` (outputs of the reference's own
``RetinaFace.call`` / ``Detection.__call__`` produced by
``oracle/make_golden.py``).
"""
import math

import numpy as np

STRIDES = (32, 16, 8)
#: (lo, hi) of the two square reference anchors per stride — the closed form of
#: generate_anchors(base 16, ratio 1, scales (32,16)/(8,4)/(2,1)).
ANCHOR_REFS = {
    32: ((-248.0, 263.0), (-120.0, 135.0)),
    16: ((-56.0, 71.0), (-24.0, 39.0)),
    8: ((-8.0, 23.0), (0.0, 15.0)),
}


def anchor_refs_from_settings(base_size, scales):
    """generate_anchors for ratio (1,) (anchors.py:75-134), to cross-check
    ANCHOR_REFS."""
    w = h = float(base_size)
    ctr = 0.5 * (w - 1)
    out = []
    for s in scales:
        ws = w * s
        out.append((ctr - 0.5 * (ws - 1), ctr + 0.5 * (ws - 1)))
    return tuple(out)


def feature_dims(H, W):
    return [(math.ceil(H / s), math.ceil(W / s)) for s in STRIDES]


def anchors_for(stride, fh, fw):
    """(fh*fw*2, 4) f32, flat index (h*fw + w)*2 + a (anchors.py:35-49)."""
    ys, xs = np.meshgrid(np.arange(fh, dtype=np.float32) * stride,
                         np.arange(fw, dtype=np.float32) * stride, indexing='ij')
    shifts = np.stack([xs, ys, xs, ys], axis=-1).reshape(-1, 1, 4)
    refs = np.array([[lo, lo, hi, hi] for lo, hi in ANCHOR_REFS[stride]], np.float32)
    return (refs[None] + shifts).reshape(-1, 4).astype(np.float32)


def exp_f32(x):
    """Correctly rounded f32 exp (evaluate in f64, round once).  torch's,
    numpy's and CUDA's f32 ``exp`` disagree with each other in the last ulp
    on ~40 % of inputs, so the oracle fixes the one definition every
    implementation can reproduce bit-for-bit; it is within 1 ulp of the
    reference's ``torch.exp`` (wrapper.py:50-51)."""
    return np.exp(np.asarray(x, np.float64)).astype(np.float32)


def decode(heads, H, W):
    """heads: the 9 model outputs (s32,s16,s8) x (prob (N,4,h,w), bbox (N,8,h,w),
    lmk (N,20,h,w)) as f32 arrays.  Returns scores (N,A), boxes (N,A,4),
    landmarks (N,A,5,2), all f32, strides concatenated 32 -> 16 -> 8."""
    heads = [np.asarray(h, dtype=np.float32) for h in heads]
    one, half = np.float32(1.0), np.float32(0.5)
    S, B, L = [], [], []
    for si, stride in enumerate(STRIDES):
        prob, bbox, lmk = heads[3 * si:3 * si + 3]
        n, _, fh, fw = prob.shape
        assert (fh, fw) == (math.ceil(H / stride), math.ceil(W / stride))
        anc = anchors_for(stride, fh, fw)
        score = prob[:, 2:].transpose(0, 2, 3, 1).reshape(n, -1)
        d = bbox.transpose(0, 2, 3, 1).reshape(n, -1, 4)
        m = lmk.transpose(0, 2, 3, 1).reshape(n, -1, 5, 2)
        aw = anc[:, 2] - anc[:, 0] + one
        ah = anc[:, 3] - anc[:, 1] + one
        cx = anc[:, 0] + half * (aw - one)
        cy = anc[:, 1] + half * (ah - one)
        pcx = d[..., 0] * aw + cx
        pcy = d[..., 1] * ah + cy
        pw = exp_f32(d[..., 2]) * aw
        ph = exp_f32(d[..., 3]) * ah
        box = np.stack([pcx - half * (pw - one), pcy - half * (ph - one),
                        pcx + half * (pw - one), pcy + half * (ph - one)], axis=-1)
        pts = np.empty_like(m)
        pts[..., 0] = m[..., 0] * aw[None, :, None] + cx[None, :, None]
        pts[..., 1] = m[..., 1] * ah[None, :, None] + cy[None, :, None]
        S.append(score), B.append(box.astype(np.float32)), L.append(pts)
    return np.concatenate(S, 1), np.concatenate(B, 1), np.concatenate(L, 1)


def nms(boxes, thr):
    """Greedy NMS over boxes already in score-descending order; returns kept
    positions (ascending = score-descending)."""
    n = boxes.shape[0]
    x1, y1, x2, y2 = (boxes[:, i].astype(np.float32) for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        j = slice(i + 1, n)
        iw = np.maximum(zero, np.minimum(x2[i], x2[j]) - np.maximum(x1[i], x1[j]))
        ih = np.maximum(zero, np.minimum(y2[i], y2[j]) - np.maximum(y1[i], y1[j]))
        inter = (iw * ih).astype(np.float32)
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / (area[i] + area[j] - inter)
        dead[j] |= iou.astype(np.float64) > float(thr)
    return np.asarray(keep, dtype=np.int64)


def select(scores, boxes, lmks, threshold=0.5, nms_threshold=0.4):
    """Per image: candidates (score >= threshold, ascending anchor index),
    stable descending sort (lower anchor index first among equal scores — the
    reference's argsort is unstable, so parity inputs must be tie-free), NMS.
    Returns a list of dicts with the surviving anchor indices and values."""
    out = []
    for s, b, l in zip(scores, boxes, lmks):
        cand = np.flatnonzero(s >= np.float32(threshold))
        order = cand[np.argsort(-s[cand], kind='stable')]
        keep = nms(b[order], nms_threshold) if len(order) else np.zeros(0, np.int64)
        idx = order[keep]
        out.append({'index': idx.astype(np.int64), 'score': s[idx], 'bbox': b[idx],
                    'landmarks': l[idx], 'num_candidates': int(len(cand))})
    return out


def model_call(heads, H, W, threshold=0.5, nms_threshold=0.4):
    """Equivalent of ``RetinaFace.call`` after the network forward."""
    s, b, l = decode(heads, H, W)
    res = select(s, b, l, threshold, nms_threshold)
    return [
        [{'bbox': r['bbox'][i], 'landmarks': r['landmarks'][i], 'score': r['score'][i]}
         for i in range(len(r['index']))]
        for r in res
    ]


def resize_out(faces_per_image, scales):
    """Detection.resize_out (detection/__init__.py:59-84): np.around (half to
    even) of f32 value / python-float scale, cast to int32."""
    if not isinstance(scales, list):
        scales = [scales] * len(faces_per_image)
    out = []
    for faces, scale in zip(faces_per_image, scales):
        out.append([{
            'bbox': np.around(f['bbox'] / scale).astype(np.int32),
            'landmarks': np.around(f['landmarks'] / scale).astype(np.int32),
            'score': f['score'],
        } for f in faces])
    return out
