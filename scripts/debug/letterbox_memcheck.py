"""tr_face_letterbox through the C ABI alone (no model), for compute-sanitizer memcheck:
images of awkward sizes, compared with the reference's PIL calls."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle.letterbox import preprocess_face_no_landmarks  # noqa: E402
from terran_b200 import _native as nat  # noqa: E402

nat.init(0)
rng = np.random.default_rng(4)
shapes = [(1, 1), (3, 5), (57, 41), (300, 181), (9, 640), (640, 9), (113, 111), (1080, 1920), (2000, 37)]
images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
n, S = len(images), 112
sizes = np.array([im.shape[:2] for im in images], np.int32)
nbytes = np.array([im.size for im in images], np.int64)
offsets = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
pixels = torch.from_numpy(np.concatenate([im.reshape(-1) for im in images])).cuda()
need = nat.lib().tr_face_letterbox_workspace_bytes(sizes.ctypes.data, n, S)
ws = torch.empty(need, dtype=torch.uint8, device='cuda')          # exactly the advertised size
out = torch.empty((n, 3, S, S), dtype=torch.uint8, device='cuda')
nat.check(nat.lib().tr_face_letterbox(C.c_void_p(pixels.data_ptr()), offsets.ctypes.data, sizes.ctypes.data,
                                      n, S, C.c_void_p(ws.data_ptr()), C.c_void_p(out.data_ptr()), None))
torch.cuda.synchronize()
got = out.cpu().numpy()
for im, g in zip(images, got):
    np.testing.assert_array_equal(g, preprocess_face_no_landmarks(im))
print('letterbox: %d images bit-exact, workspace %d bytes' % (n, need))
