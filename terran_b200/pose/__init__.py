"""``Estimation`` task wrapper, the ``Keypoint`` enum and ``pose_estimation``.

Drop-in for the reference's ``terran/pose/__init__.py`` (``Keypoint`` :13-36,
``Estimation`` :131-223): one HWC image, an (N,H,W,3) array or a list of images
in; per image a list of ``{'keypoints': (18,3) int (x, y, present), 'score'}``
out.  Images are NOT resized here — the model class resizes to ``short_side``
itself, as upstream.
"""
from enum import Enum

import numpy as np
import torch

from terran_b200.batching import PadMerge
from terran_b200.checkpoint import get_class_for_checkpoint
from terran_b200.defaults import default_device

TASK_NAME = 'pose-estimation'


class Keypoint(Enum):
    NOSE = 0
    NECK = 1
    R_SHOULDER = 2
    R_ELBOW = 3
    R_HAND = 4
    L_SHOULDER = 5
    L_ELBOW = 6
    L_HAND = 7
    R_HIP = 8
    R_KNEE = 9
    R_FOOT = 10
    L_HIP = 11
    L_KNEE = 12
    L_FOOT = 13
    R_EYE = 14
    L_EYE = 15
    R_EAR = 16
    L_EAR = 17


class Estimation:

    def __init__(self, checkpoint=None, short_side=184, merge_method='padding',
                 device=default_device, lazy=False):
        """short_side defaults to 184 "to keep the model fast enough" upstream
        (386 for better results); other kwargs as in ``Detection``."""
        self.device = device
        self.estimation_cls = get_class_for_checkpoint(TASK_NAME, checkpoint)
        self.short_side = short_side
        self.model = (
            self.estimation_cls(device=self.device, short_side=self.short_side)
            if not lazy else None)
        self.merger = PadMerge(merge_method)

    def __repr__(self):
        return f'<Estimation({self.estimation_cls.__name__})>'

    def submit(self, images):
        """Start pose estimation and return a handle with ``result()`` (see
        ``Detection.submit``): for array / CUDA batches the work is only enqueued."""
        from terran_b200.face.detection import _Deferred, _first
        single = not isinstance(images, (list, tuple)) and len(images.shape) == 3
        if single:
            images = images[None] if isinstance(images, torch.Tensor) else np.expand_dims(images, 0)
        batch, offsets = self.merger.merge(images)
        if self.model is None:
            self.model = self.estimation_cls(device=self.device, short_side=self.short_side)
        model = self.model
        if offsets is None and hasattr(model, 'estimate_async'):
            from terran_b200.frames import to_device_u8
            with torch.cuda.device(model.device_index):
                pending = model.estimate_async(to_device_u8(batch, model.device_index))
            return _Deferred(lambda: _first(pending.result(), single))
        poses = self.merger.unpad_poses(model.call(batch), offsets)
        return _Deferred(lambda: _first(poses, single))

    def __call__(self, images):
        return self.submit(images).result()


pose_estimation = Estimation(lazy=True)
"""Default entry point to pose estimation (lazily loaded, reference :226)."""
