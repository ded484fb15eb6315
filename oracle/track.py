"""ORACLE (test infrastructure, not product code): loop-form restatement of the reference's
SORT face tracker (``terran/tracking/face.py:9-411``) and of the Kalman filter it takes from
``filterpy`` (an un-pinned dependency in the reference's ``setup.py``; absent from this
image, so PARITY OF THE FILTER ARITHMETIC IS UNPINNED: ``LinearKalman`` restates filterpy
1.4.5's published ``KalmanFilter.predict/update`` — x = Fx, P = FPF' + Q; y = z - Hx,
S = HPH' + R, K = PH'S^-1, x += Ky, P = (I-KH)P(I-KH)' + KRK').  The ASSOCIATION logic (IoU
matrix, Hungarian assignment, confirmation rules, output order, id numbering) is pinned by
running the reference's own ``Sort`` with ``filterpy.kalman.KalmanFilter`` stubbed by
``LinearKalman`` (``oracle/make_golden_track.py`` -> ``tests/golden/sort_tracking.npz``).

Only ``tests/`` may import this.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


class LinearKalman:
    """filterpy.kalman.KalmanFilter (defaults: x = 0, P = Q = R = F = I, H = 0)."""

    def __init__(self, dim_x, dim_z):
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self._I = np.eye(dim_x)

    def predict(self):
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + self.Q

    def update(self, z):
        y = z - self.H @ self.x
        PHT = self.P @ self.H.T
        S = self.H @ PHT + self.R
        K = PHT @ np.linalg.inv(S)
        self.x = self.x + K @ y
        I_KH = self._I - K @ self.H
        self.P = I_KH @ self.P @ I_KH.T + K @ self.R @ K.T


def iou(a, b):
    """face.py:14-44 (areas without +1)."""
    w = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    h = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = w * h
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter)


def to_center(b):
    """face.py:47-70: (x_min, y_min, x_max, y_max) -> (cx, cy, area, ratio) column."""
    w, h = b[2] - b[0], b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / h], dtype=np.float64).reshape(4, 1)


def to_corners(x):
    """face.py:73-94."""
    w = np.sqrt(x[2] * x[3])
    h = x[2] / w
    return np.concatenate([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0])


class Track:
    """face.py:97-193 (KalmanTracker).  Ids are numbered by the owning SortOracle."""

    def __init__(self, bbox, track_id):
        kf = LinearKalman(7, 4)
        kf.F[0, 4] = kf.F[1, 5] = kf.F[2, 6] = 1.0
        kf.H[:4, :4] = np.eye(4)
        kf.R[2:, 2:] *= 10.0
        kf.P[4:, 4:] *= 1000.0
        kf.P *= 10.0
        kf.Q[-1, -1] *= 0.01
        kf.Q[4:, 4:] *= 0.01
        kf.x[:4] = to_center(bbox)
        self.kf, self.hits, self.age, self.id = kf, 0, 0, track_id

    def predict(self):
        if self.kf.x[6] + self.kf.x[2] <= 0:
            self.kf.x[6] *= 0.0
        self.kf.predict()
        self.age += 1
        return to_corners(self.kf.x)

    def update(self, bbox):
        self.age = 0
        self.hits += 1
        self.kf.update(to_center(bbox))


class SortOracle:
    """face.py:196-411 (associate_detections_to_trackers + Sort.update)."""

    def __init__(self, max_age=1, min_hits=3, return_unmatched=False, iou_threshold=0.3):
        self.max_age, self.min_hits, self.return_unmatched = max_age, min_hits, return_unmatched
        self.iou_threshold = iou_threshold
        self.tracks, self.frames, self.next_id = [], 0, 0

    def update(self, faces):
        self.frames += 1
        boxes = [t.predict() for t in self.tracks]
        keep = [i for i, b in enumerate(boxes) if not np.any(np.isnan(b))]
        self.tracks = [self.tracks[i] for i in keep]
        boxes = [boxes[i] for i in keep]

        F, T = len(faces), len(boxes)
        matches, lone_faces, lone_tracks = [], [], []
        if T == 0:
            lone_faces = list(range(F))
        else:
            m = np.zeros((F, T), dtype=np.float32)
            for f in range(F):
                for t in range(T):
                    m[f, t] = iou(faces[f]['bbox'], boxes[t])
            rows, cols = linear_sum_assignment(-m)
            lone_faces = [f for f in range(F) if f not in rows]
            lone_tracks = [t for t in range(T) if t not in cols]
            for f, t in zip(rows, cols):
                if m[f, t] < self.iou_threshold:
                    lone_faces.append(f)
                    lone_tracks.append(t)
                else:
                    matches.append((f, t))

        out = []
        for t, track in enumerate(self.tracks):
            if t in lone_tracks:
                continue
            f = next(ff for ff, tt in matches if tt == t)
            track.update(faces[f]['bbox'])
            confirmed = track.hits >= self.min_hits or self.frames <= self.min_hits
            out.append({'track': track.id if confirmed else None, **faces[f]})
        for f in lone_faces:
            track = Track(faces[f]['bbox'], self.next_id)
            self.next_id += 1
            self.tracks.append(track)
            out.append({'track': track.id if self.min_hits == 0 else None, **faces[f]})
        if not self.return_unmatched:
            out = [o for o in out if o['track'] is not None]
        self.tracks = [t for t in self.tracks if t.age <= self.max_age]
        return out


def synthetic_sequence(seed, frames=40, people=4, size=(1080, 1920), jitter=3.0, drop=0.15):
    """Boxes of `people` faces drifting linearly with jitter; detections are dropped at random
    and shuffled; one person enters late and one leaves early.  Returns a list (per frame) of
    lists of face dicts with int32 bbox (as `Detection` returns them) and a score."""
    rng = np.random.default_rng(seed)
    H, W = size
    pos = rng.uniform([200, 200], [W - 200, H - 200], (people, 2))
    vel = rng.uniform(-12, 12, (people, 2))
    side = rng.uniform(60, 180, people)
    first = np.zeros(people, int)
    last = np.full(people, frames)
    if people > 1:
        first[-1] = frames // 3
        last[0] = 2 * frames // 3
    out = []
    for t in range(frames):
        faces = []
        for p in range(people):
            if not (first[p] <= t < last[p]) or rng.random() < drop:
                continue
            c = pos[p] + vel[p] * t + rng.normal(0, jitter, 2)
            s = side[p] * (1 + 0.01 * t) + rng.normal(0, jitter)
            box = np.array([c[0] - s / 2, c[1] - s * 0.6, c[0] + s / 2, c[1] + s * 0.6])
            faces.append({'bbox': np.around(box).astype(np.int32), 'score': np.float32(rng.uniform(0.6, 1.0)),
                          'person': p})
        order = rng.permutation(len(faces))
        out.append([faces[i] for i in order])
    return out
