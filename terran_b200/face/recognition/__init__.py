"""``Recognition`` task wrapper and the ``extract_features`` entry point.

Drop-in for the reference's ``terran/face/recognition/__init__.py:7-93``: a
single HWC image (optionally with one face dict or a list of them), or a list
of images (optionally with a list of face lists); raises ``ValueError`` when
``images`` and ``faces_per_image`` disagree in length (:73-78).
"""
from terran_b200.checkpoint import get_class_for_checkpoint
from terran_b200.defaults import default_device

TASK_NAME = 'face-recognition'


class Recognition:

    def __init__(self, checkpoint=None, device=default_device, lazy=False):
        self.device = device
        self.recognition_cls = get_class_for_checkpoint(TASK_NAME, checkpoint)
        self.model = self.recognition_cls(device=self.device) if not lazy else None

    def __repr__(self):
        return f'<Recognition({self.recognition_cls.__name__})>'

    def __call__(self, images, faces_per_image=None):
        """Returns one (N_i, 512) float32 array per image (or a single array /
        vector when a single image / single face was passed)."""
        single = not isinstance(images, (list, tuple)) and len(images.shape) == 3
        one_face = isinstance(faces_per_image, dict)
        if single:
            images = [images]
            faces_per_image = [[faces_per_image]] if one_face else [faces_per_image]

        if faces_per_image is not None and len(faces_per_image) != len(images):
            raise ValueError(
                f'`images` and `faces_per_image` must be of the same size, '
                f'but the former is of size {len(images)} while the latter of '
                f'size {len(faces_per_image)}.'
            )

        if self.model is None:
            self.model = self.recognition_cls(device=self.device)
        out = self.model.call(images, faces_per_image)

        if single and one_face:
            return out[0][0]
        return out[0] if single else out


extract_features = Recognition(lazy=True)
"""Default entry point to face recognition (lazily loaded, reference :93)."""
