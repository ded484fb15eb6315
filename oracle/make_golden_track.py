"""ORACLE tooling: run the reference's OWN ``Sort`` (``/root/reference/terran/tracking/face.py``,
build container only) on seeded synthetic detection sequences and write
``tests/golden/sort_tracking.npz``.  ``filterpy`` is not installed: the module is stubbed with
``oracle.track.LinearKalman`` (a restatement of filterpy's published equations — that part of the
parity is therefore unpinned); everything else, i.e. the association, confirmation, ordering and
id rules, is the reference's code.

    python oracle/make_golden_track.py
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get('TERRAN_REFERENCE', '/root/reference')

import numpy as np  # noqa: E402

from oracle import track  # noqa: E402

CASES = [dict(seed=s, frames=40, people=p, max_age=a, min_hits=h, return_unmatched=r)
         for s, p, a, h, r in [(0, 4, 1, 3, False), (1, 6, 5, 2, True), (2, 3, 30, 6, False),
                               (3, 8, 2, 0, False), (4, 1, 1, 3, True), (5, 5, 3, 1, False)]]


def reference_sort_module():
    fp = types.ModuleType('filterpy')
    fk = types.ModuleType('filterpy.kalman')
    fk.KalmanFilter = lambda dim_x, dim_z: track.LinearKalman(dim_x, dim_z)
    sys.modules['filterpy'], sys.modules['filterpy.kalman'] = fp, fk
    # terran.tracking.face imports terran.face.detection (-> the whole package) only for the
    # FaceTracking wrapper; give it a stand-in so that the file loads on its own.
    det = types.ModuleType('terran.face.detection')
    det.Detection, det.face_detection = type('Detection', (), {}), None
    for name in ('terran', 'terran.face'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['terran.face.detection'] = det
    spec = importlib.util.spec_from_file_location('ref_tracking_face',
                                                  os.path.join(REFERENCE, 'terran', 'tracking', 'face.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def encode(per_frame):
    """[(frame, person, track or -1, x1, y1, x2, y2)] in output order."""
    rows = []
    for t, faces in enumerate(per_frame):
        for f in faces:
            rows.append([t, f['person'], -1 if f['track'] is None else f['track'], *f['bbox']])
    return np.array(rows, dtype=np.int64).reshape(-1, 7)


def main():
    ref = reference_sort_module()
    out = {}
    for i, c in enumerate(CASES):
        seq = track.synthetic_sequence(c['seed'], c['frames'], c['people'])
        ref.KalmanTracker.count = 0                      # ids are a process-wide counter there
        s = ref.Sort(max_age=c['max_age'], min_hits=c['min_hits'], return_unmatched=c['return_unmatched'])
        got = [s.update(faces) for faces in seq]
        o = track.SortOracle(c['max_age'], c['min_hits'], c['return_unmatched'])
        mine = [o.update(faces) for faces in seq]
        a, b = encode(got), encode(mine)
        print(f'case {i} {c}: reference rows {len(a)}, oracle identical: {np.array_equal(a, b)}')
        out[f'case{i}'] = a
        out[f'cfg{i}'] = np.array([c['seed'], c['frames'], c['people'], c['max_age'], c['min_hits'],
                                   int(c['return_unmatched'])])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'sort_tracking.npz'), **out)


if __name__ == '__main__':
    main()
