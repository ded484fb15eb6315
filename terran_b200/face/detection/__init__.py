"""``Detection`` task wrapper and the ``face_detection`` entry point.

Drop-in for the reference's ``terran/face/detection/__init__.py`` (``Detection``
:185-287): same constructor kwargs, same accepted inputs (one HWC image, an
(N,H,W,3) array, or a list of differently sized images), same output dicts —
``bbox`` int32 (4,), ``landmarks`` int32 (5,2), ``score`` float32 — and the same
errors.

One difference in mechanism, none in result: an (N,H,W,3) batch (or a CUDA
uint8 tensor) is resized on the GPU by a kernel that reproduces
``cv2.resize(INTER_LINEAR)`` bit for bit, instead of a per-image host loop.
Lists of images of different sizes take the host resize + pad-merge path.
"""
import numpy as np
import torch

from terran_b200.batching import PadMerge, host_resize, round_faces
from terran_b200.checkpoint import get_class_for_checkpoint
from terran_b200.defaults import cuda_index, default_device
from terran_b200.frames import resize_short_side, to_device_u8

TASK_NAME = 'face-detection'


class Detection:

    def __init__(self, checkpoint=None, short_side=416, merge_method='padding',
                 device=default_device, lazy=False):
        """checkpoint : alias of the model to use (``None`` = task default).
        short_side : images are resized so their short side has this length.
        merge_method : 'padding' (or 'crop', not implemented — as upstream).
        device : torch device holding the model.  lazy : defer model loading."""
        self.device = device
        self.short_side = short_side
        self.detection_cls = get_class_for_checkpoint(TASK_NAME, checkpoint)
        self.model = self.detection_cls(device=self.device) if not lazy else None
        self.merger = PadMerge(merge_method)
        #: resize array batches on the GPU (bit-exact with cv2); False forces the
        #: host cv2 loop.
        self.device_resize = True

    def __repr__(self):
        return f'<Detection({self.detection_cls.__name__})>'

    def _model(self):
        if self.model is None:
            self.model = self.detection_cls(device=self.device)
        return self.model

    def submit(self, images):
        """Start face detection on ``images`` and return a handle whose
        ``result()`` gives what ``__call__`` returns.  For array / CUDA-tensor
        batches everything up to the result download is only ENQUEUED on the
        current CUDA stream (no host synchronisation), so the caller can overlap
        the next batch or other work; lists of images are processed right away."""
        single = not isinstance(images, (list, tuple)) and len(images.shape) == 3
        if single:
            images = images[None]

        if isinstance(images, (np.ndarray, torch.Tensor)) and self.device_resize:
            model = self._model()
            idx = cuda_index(self.device)
            with torch.cuda.device(idx):
                frames, scales = resize_short_side(to_device_u8(images, idx), self.short_side)
                if hasattr(model, 'detect_async'):
                    pending = model.detect_async(frames)
                    handle = _Deferred(lambda: _first(pending.result(scale=scales), single))
                    handle.pending, handle.scale = pending, scales     # for device-resident consumers
                    return handle
            offsets = None
        elif isinstance(images, (list, tuple)) and self.device_resize and len(images):
            # A list of (possibly differently sized) images: every image is uploaded and resized
            # on the device (bit-exact with the reference's host cv2.resize), and the centred
            # zero-pad merge (detection/__init__.py:86-137) is built in device memory.
            model = self._model()
            idx = cuda_index(self.device)
            with torch.cuda.device(idx):
                resized, scales = [], []
                for img in images:
                    img = img if isinstance(img, torch.Tensor) else np.asarray(img)
                    t, s = resize_short_side(to_device_u8(img[None], idx), self.short_side)
                    resized.append(t[0])
                    scales.append(s)
                frames, offsets = self.merger.merge_device(resized)
        else:
            if isinstance(images, torch.Tensor):
                images = images.cpu().numpy()
            resized, scales = host_resize(images, self.short_side)
            frames, offsets = self.merger.merge(resized)
            model = self._model()

        faces = model.call(frames)
        faces = self.merger.unpad_faces(faces, offsets)
        faces = round_faces(faces, scales)
        return _Deferred(lambda: _first(faces, single))

    def __call__(self, images):
        """Face detection on one image, an array batch or a list of images.
        Returns a list of face dicts per image (or one list for one image)."""
        return self.submit(images).result()


def _first(out, single):
    return out[0] if single else out


class _Deferred:
    """Handle returned by ``submit``: ``result()`` finishes the computation once."""

    def __init__(self, finish):
        self._finish, self._done, self._value = finish, False, None

    def result(self):
        if not self._done:
            self._value, self._done, self._finish = self._finish(), True, None
        return self._value


face_detection = Detection(lazy=True)
"""Default entry point to face detection — lazily loaded, like the reference's
``face_detection`` (``terran/face/detection/__init__.py:290``)."""
