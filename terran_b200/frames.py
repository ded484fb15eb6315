"""Host<->device frame plumbing shared by the three model wrappers: pinned
staging of uint8 frames, the on-device cv2-exact resize and result download.

``resize_short_side`` replaces the host ``cv2.resize(INTER_LINEAR)`` loops of
the reference (``terran/face/detection/__init__.py:15-57``,
``terran/pose/openpose/wrapper.py:93-113``) with one kernel over the whole
batch; sizes and scale follow the reference exactly
(``scale = short_side / min(H, W)``, ``dsize = (int(W*scale), int(H*scale))``).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat


def to_device_u8(images, device_index):
    """(N,H,W,3) uint8 numpy array or torch tensor -> contiguous CUDA tensor."""
    if isinstance(images, torch.Tensor):
        t = images
        if t.dtype != torch.uint8:
            raise TypeError('frames must be uint8')
        if t.device.type != 'cuda':
            t = t.pin_memory().to(f'cuda:{device_index}', non_blocking=True)
        return t.contiguous()
    arr = np.ascontiguousarray(images)
    if arr.dtype != np.uint8:
        # The reference casts whatever it gets to float32; its documented input
        # is uint8 RGB, which is the only dtype the stem kernels read.
        arr = arr.astype(np.uint8)
    return torch.from_numpy(arr).pin_memory().to(f'cuda:{device_index}', non_blocking=True)


def resized_dims(H, W, short_side):
    scale = short_side / min(H, W)
    return int(H * scale), int(W * scale), scale


def resize_short_side(frames, short_side):
    """frames: CUDA uint8 (N,H,W,3).  Returns (resized CUDA uint8, scale)."""
    N, H, W, _ = frames.shape
    h, w, scale = resized_dims(H, W, short_side)
    if (h, w) == (H, W):
        return frames, scale
    out = torch.empty((N, h, w, 3), dtype=torch.uint8, device=frames.device)
    nat.check(nat.lib().tr_resize_bilinear_u8(
        C.c_void_p(frames.data_ptr()), N, H, W, C.c_void_p(out.data_ptr()), h, w,
        nat.current_stream_ptr()))
    return out, scale
