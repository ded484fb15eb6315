"""ORACLE (test infrastructure, not product code): integer restatement of
``cv2.resize(src, dsize, interpolation=INTER_LINEAR)`` on uint8 images — the
host resize the reference runs in ``terran/face/detection/__init__.py:15-57``
and ``terran/pose/openpose/wrapper.py:93-113`` (cv2 itself is a third-party
dependency outside the reference tree; this follows OpenCV 4.x
``resizeGeneric_`` with ``HResizeLinear`` / ``VResizeLinear`` for 8-bit data,
SURVEY.md Appendix A.5).  Pinned against the installed cv2 by
``tests/test_oracle_golden.py::test_resize_oracle_matches_cv2``.
"""
import numpy as np


def _taps(n_dst, n_src, clamp_weights):
    """Per output index: (i0, i1, a0, a1) — source taps and 11-bit weights.
    OpenCV treats the two axes differently at the borders: along x a tap that
    falls outside the image is clamped AND its fraction reset to 0; along y only
    the row INDEX is clipped when the rows are fetched, the weights keep the
    unclamped fraction (visible when up-scaling: both taps hit the border row
    and the two truncated products can sum to one less)."""
    inv_scale = n_dst / n_src            # double, like OpenCV's inv_scale_x
    scale = 1.0 / inv_scale
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        s[lo], f[lo] = 0, 0
        hi = s >= n_src - 1
        s[hi], f[hi] = n_src - 1, 0
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1), a0, a1


def resize_bilinear_u8(img, h, w):
    """img (H,W,C) uint8 -> (h,w,C) uint8."""
    H, W = img.shape[:2]
    x0, x1, ax0, ax1 = _taps(w, W, True)
    y0, y1, ay0, ay1 = _taps(h, H, False)
    src = img.astype(np.int64)
    rows = src[:, x0] * ax0[None, :, None] + src[:, x1] * ax1[None, :, None]     # (H,w,C)
    top, bot = rows[y0] >> 4, rows[y1] >> 4
    out = (((ay0[:, None, None] * top) >> 16) + ((ay1[:, None, None] * bot) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_short_side(img, short_side):
    H, W = img.shape[:2]
    scale = short_side / min(H, W)
    return resize_bilinear_u8(img, int(H * scale), int(W * scale)), scale
