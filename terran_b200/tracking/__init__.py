"""Face tracking on detections (SORT): drop-in for ``terran/tracking/__init__.py``."""
from terran_b200.tracking.face import FaceTracking, Sort, face_tracking  # noqa
