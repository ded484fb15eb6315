"""Child process of tests/test_reference_plugin.py: import the UNMODIFIED reference
(baseline/_ref or /root/reference) with the shims of baseline/refarm.py, plug the B200 classes in
under the alias 'b200' and drive them through the reference's own wrappers."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode = sys.argv[1]
if mode == 'cpu':
    os.environ['CUDA_VISIBLE_DEVICES'] = ''

import numpy as np  # noqa: E402
import torch  # noqa: E402

from baseline import refarm  # noqa: E402
from terran_b200 import synth  # noqa: E402

terran = refarm.import_reference({'openpose': synth.openpose_state_dict(peaks=True)})
from terran_b200.checkpoint import register_with_reference  # noqa: E402

register_with_reference()
register_with_reference()        # idempotent
import terran.face  # noqa: E402
import terran.pose  # noqa: E402
from terran.checkpoint import CHECKPOINTS  # noqa: E402

out = {'n_b200_entries': sum(c['alias'] == 'b200' for c in CHECKPOINTS)}
det = terran.face.Detection(checkpoint='b200', lazy=True)
rec = terran.face.Recognition(checkpoint='b200', lazy=True)
est = terran.pose.Estimation(checkpoint='b200', lazy=True)
out['classes'] = [f'{c.__module__}.{c.__name__}' for c in
                  (det.detection_cls, rec.recognition_cls, est.estimation_cls)]
out['defaults'] = [f'{c.__module__}.{c.__name__}' for c in
                   (terran.face.Detection(lazy=True).detection_cls,
                    terran.face.Recognition(lazy=True).recognition_cls,
                    terran.pose.Estimation(lazy=True).estimation_cls)]
out['repr'] = repr(det)
try:
    terran.face.Detection(checkpoint='no-such-alias', lazy=True)
    out['bad_alias'] = 'no error'
except ValueError as e:
    out['bad_alias'] = str(e)

if mode == 'gpu':
    # the reference's wrappers (host cv2 resize, pad-merge, rounding) around the B200 classes
    from terran_b200.face.detection import Detection as OurDetection
    from terran_b200.pose import Estimation as OurEstimation
    frames = np.random.default_rng(5).integers(0, 256, (3, 540, 960, 3), dtype=np.uint8)
    dev = torch.device('cuda')
    ref_det = terran.face.Detection(checkpoint='b200', device=dev)
    ref_est = terran.pose.Estimation(checkpoint='b200', device=dev)
    ours_det, ours_est = OurDetection(device=dev), OurEstimation(device=dev)
    a, b = ref_det(frames), ours_det(frames)
    out['faces'] = [len(f) for f in a]
    out['faces_equal'] = all(
        len(x) == len(y) and all(np.array_equal(p['bbox'], q['bbox']) and
                                 np.array_equal(p['landmarks'], q['landmarks']) and
                                 p['score'] == q['score'] and p['bbox'].dtype == q['bbox'].dtype
                                 for p, q in zip(x, y)) for x, y in zip(a, b))
    a, b = ref_est(frames), ours_est(frames)
    out['humans'] = [len(f) for f in a]
    out['humans_equal'] = all(
        len(x) == len(y) and all(np.array_equal(p['keypoints'], q['keypoints']) and
                                 p['score'] == q['score'] for p, q in zip(x, y))
        for x, y in zip(a, b))
    # single image and list-of-different-sizes inputs go through the reference's own merge code
    one = ref_det(frames[0])
    out['single_equal'] = len(one) == len(ours_det(frames[0]))
    lst = ref_det([frames[0], frames[1][:400, :700]])
    out['list_lens'] = [len(f) for f in lst]
    crops = np.random.default_rng(6).integers(0, 256, (4, 112, 112, 3), dtype=np.uint8)
    ref_rec = terran.face.Recognition(checkpoint='b200', device=dev)
    emb = ref_rec(list(crops))
    out['emb_shape'] = list(np.asarray(emb).shape)
    out['emb_norm_ok'] = bool(np.allclose(np.linalg.norm(np.asarray(emb), axis=1), 1.0, atol=1e-5))
print('RESULT ' + json.dumps(out))
