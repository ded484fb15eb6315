"""Host-side batching shared by ``Detection`` and ``Estimation``: short-side
resize of image lists and centred zero-pad merging, with the inverse
coordinate fix-ups.

Behavioural contract (what the reference does in
``terran/face/detection/__init__.py:13-182`` and ``terran/pose/__init__.py:41-128``):
  * scale = short_side / min(H, W); dsize = (int(W*scale), int(H*scale)), cv2 bilinear;
  * a list of images is padded to the largest height/width with the extra rows
    split ceil (top/left) / floor (bottom/right);
  * an ndarray batch passes through untouched, even when ``method`` is invalid;
  * ``method='crop'`` raises NotImplementedError, anything else ValueError — but
    only when a list actually has to be merged.
"""
import numpy as np
import torch


def host_resize(images, short_side):
    """cv2 resize of an (N,H,W,3) array (one scale) or a list (one scale each)."""
    import cv2

    def one(img):
        h, w = img.shape[:2]
        scale = short_side / min(h, w)
        size = (int(w * scale), int(h * scale))
        return cv2.resize(src=img, dsize=size, interpolation=cv2.INTER_LINEAR), scale

    if isinstance(images, np.ndarray):
        outs = [one(img) for img in images]
        return np.stack([o[0] for o in outs]), outs[0][1]
    outs = [one(img) for img in images]
    return [o[0] for o in outs], [o[1] for o in outs]


class PadMerge:
    """Zero-pad merge of a list of HWC uint8 images into one batch."""

    def __init__(self, method='padding'):
        self.method = method

    def _check(self):
        if self.method == 'crop':
            raise NotImplementedError
        if self.method != 'padding':
            raise ValueError('Invalid `method` set, options are `padding` or `crop`.')

    def merge(self, images):
        """Returns (batch, offsets) where offsets is None for an already merged
        batch, else an (N,2) int array of (left, top) pads."""
        if isinstance(images, (np.ndarray, torch.Tensor)):
            return images, None
        self._check()
        hs = np.array([im.shape[0] for im in images])
        ws = np.array([im.shape[1] for im in images])
        H, W = int(hs.max()), int(ws.max())
        top = -((hs - H) // 2)          # ceil((H - h) / 2)
        left = -((ws - W) // 2)
        batch = np.zeros((len(images), H, W, 3), np.uint8)
        for i, im in enumerate(images):
            batch[i, top[i]:top[i] + hs[i], left[i]:left[i] + ws[i]] = im
        return batch, np.stack([left, top], axis=1)

    def merge_device(self, images):
        """``merge`` for a list of CUDA uint8 (h,w,3) tensors (already resized on the device):
        the same centred zero padding, built in device memory — no host copy of the pixels."""
        self._check()
        hs = np.array([int(im.shape[0]) for im in images])
        ws = np.array([int(im.shape[1]) for im in images])
        H, W = int(hs.max()), int(ws.max())
        top = -((hs - H) // 2)
        left = -((ws - W) // 2)
        batch = torch.zeros((len(images), H, W, 3), dtype=torch.uint8, device=images[0].device)
        for i, im in enumerate(images):
            batch[i, top[i]:top[i] + hs[i], left[i]:left[i] + ws[i]] = im
        return batch, np.stack([left, top], axis=1)

    def unpad_faces(self, faces_per_image, offsets):
        if offsets is None:
            return faces_per_image
        self._check()
        out = []
        for faces, off in zip(faces_per_image, offsets):
            out.append([
                # dtypes follow the reference: bbox stays float32 (scalar - int),
                # landmarks promote to float64 (float32 array - int64 array)
                {'bbox': f['bbox'] - np.tile(off, 2).astype(np.float32),
                 'landmarks': f['landmarks'] - off[None, :], 'score': f['score']}
                for f in faces
            ])
        return out

    def unpad_poses(self, poses_per_image, offsets):
        if offsets is None:
            return poses_per_image
        self._check()
        out = []
        for poses, off in zip(poses_per_image, offsets):
            shift = np.array([off[0], off[1], 0]).reshape(1, 3)
            fixed = []
            for pose in poses:
                kp = pose['keypoints'] - shift
                kp[kp[..., 2] == 0] = 0          # absent joints stay (0, 0, 0)
                fixed.append({'keypoints': kp, 'score': pose['score']})
            out.append(fixed)
        return out


def round_faces(faces_per_image, scales):
    """Map detections back to input-image pixels: round-half-even of
    coordinate / scale, int32 (reference ``resize_out`` :59-84)."""
    if not isinstance(scales, list):
        scales = [scales] * len(faces_per_image)
    return [
        [
            {'bbox': np.around(f['bbox'] / s).astype(np.int32),
             'landmarks': np.around(f['landmarks'] / s).astype(np.int32),
             'score': f['score']}
            for f in faces
        ]
        for faces, s in zip(faces_per_image, scales)
    ]
