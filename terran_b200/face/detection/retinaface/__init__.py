from terran_b200.face.detection.retinaface.wrapper import RetinaFace  # noqa
