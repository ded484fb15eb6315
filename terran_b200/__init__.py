"""terran_b200 — B200-native drop-in for Terran's per-frame perception path.

Same callables as the reference's ``terran/__init__.py:2-5``::

    from terran_b200 import face_detection, extract_features, pose_estimation
"""
# flake8: noqa
from terran_b200.defaults import default_device

from terran_b200.face import extract_features, face_detection, Detection, Recognition
from terran_b200.pose import pose_estimation, Estimation, Keypoint
