from terran_b200.face.detection import face_detection, Detection  # noqa
from terran_b200.face.recognition import extract_features, Recognition  # noqa
