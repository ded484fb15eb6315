from terran_b200.face.recognition.arcface.wrapper import ArcFace  # noqa
