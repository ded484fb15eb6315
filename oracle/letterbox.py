"""ORACLE (test infrastructure, not product code): faces given WITHOUT landmarks,
``preprocess_face_no_landmarks`` of ``terran/face/recognition/arcface/wrapper.py:75-99`` — PIL
``Image.resize`` to longer side 112 (Pillow's default filter for RGB images: the antialiased
BICUBIC resampler), centred on a zero 112x112 canvas, CHW, channels flipped to BGR.

Two forms: ``preprocess_face_no_landmarks`` makes the reference's own PIL calls (PIL is the
dependency the reference calls, present here and on the GPU box), and ``pil_resize_bicubic`` /
``resample_table`` restate the algorithm of Pillow's 8-bit resampler (``src/libImaging/
Resample.c``, Pillow 12.2.0: ``precompute_coeffs``, ``normalize_coeffs_8bpc``, the horizontal and
the vertical pass) in numpy.  The restatement is pinned against PIL itself in
``tests/test_oracle_golden.py::test_pil_resize_restatement_matches_pil`` (bit-exact), and is what the
tables of the product's ``tr_resample_table`` are compared with.  Product path:
``tr_face_letterbox`` (``terran_b200/csrc/detect_post.cu``).
"""
import math

import numpy as np
from PIL import Image

PRECISION_BITS = 32 - 8 - 2


def preprocess_face_no_landmarks(image, image_side=112):
    """The reference's calls (:75-99): (3, side, side) uint8 BGR."""
    face = Image.fromarray(image)
    scale = image_side / max(face.size[0], face.size[1])
    face = face.resize((int(face.size[0] * scale), int(face.size[1] * scale)))
    x_min = int((image_side - face.size[0]) / 2)
    y_min = int((image_side - face.size[1]) / 2)
    out = np.zeros((3, image_side, image_side), dtype=np.uint8)
    out[:, y_min:y_min + face.size[1], x_min:x_min + face.size[0]] = (
        np.asarray(face).transpose([2, 0, 1])[::-1, ...])
    return out


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_table(in_size, out_size):
    """(bounds (out,2) int32 [first input sample, count], coeffs (out,ksize) int32 with 22
    fractional bits) of one axis: ``precompute_coeffs`` + ``normalize_coeffs_8bpc``."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    coeffs = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        weights = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        total = 0.0
        for w in weights:
            total += w
        for x, w in enumerate(weights):
            if total != 0.0:
                w = w / total
            coeffs[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(
                0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs


def pil_resize_bicubic(image, out_w, out_h):
    """``Image.fromarray(image).resize((out_w, out_h))`` for (H,W,3) uint8: horizontal pass, round
    to uint8, vertical pass, round to uint8 (``ImagingResampleHorizontal_8bpc`` / ``Vertical``)."""
    h, w, _ = image.shape
    bh, kh = resample_table(w, out_w)
    bv, kv = resample_table(h, out_h)
    half = 1 << (PRECISION_BITS - 1)
    src = image.astype(np.int64)
    rows = np.zeros((h, out_w, 3), np.int64)
    for x in range(out_w):
        first, count = bh[x]
        acc = (src[:, first:first + count] * kh[x, :count, None].astype(np.int64)).sum(1) + half
        rows[:, x] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.zeros((out_h, out_w, 3), np.uint8)
    for y in range(out_h):
        first, count = bv[y]
        acc = (rows[first:first + count] * kv[y, :count, None, None].astype(np.int64)).sum(0) + half
        out[y] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def letterbox(image, image_side=112):
    """``preprocess_face_no_landmarks`` through the restated resampler."""
    h, w, _ = image.shape
    scale = image_side / max(w, h)
    ow, oh = int(w * scale), int(h * scale)
    x_min, y_min = int((image_side - ow) / 2), int((image_side - oh) / 2)
    out = np.zeros((3, image_side, image_side), dtype=np.uint8)
    out[:, y_min:y_min + oh, x_min:x_min + ow] = pil_resize_bicubic(image, ow, oh).transpose(2, 0, 1)[::-1]
    return out
