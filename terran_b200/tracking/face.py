"""SORT face tracking on gathered, frame-ordered detections.

Drop-in for the reference's ``terran/tracking/face.py`` (``Sort`` :317-411,
``associate_detections_to_trackers`` :196-273, ``KalmanTracker`` :97-193, ``FaceTracking``
:414-468, ``face_tracking`` :471-552): same entry point, same per-face output dicts with an
extra ``track`` field, same confirmation / ageing rules, same output order and id numbering.

Where it runs: the reference tracks on the host, one frame after the other, on the results of
a whole batch — the only consumer of the perception path that needs frame ORDER (SURVEY.md
section 8e/8f).  It stays on the host here too: per frame it is a handful of 7x7 products for a
few faces.  Unlike the reference, which keeps one ``filterpy`` object per face and loops over
faces x tracks in Python, all tracks of a frame live in stacked arrays — state (T,7), covariance
(T,7,7) — and predict / IoU / update are single vectorised numpy expressions.  The Kalman
filter is the textbook constant-velocity model the reference configures in ``filterpy``
(x = Fx, P = FPF' + Q; K = PH'(HPH' + R)^-1, x += K(z - Hx), P = (I-KH)P(I-KH)' + KRK').

Two upstream bugs are not reproduced: ``face_tracking()`` without a
``video`` crashes upstream (it dereferences ``video.framerate`` unconditionally, :546-550) and
``FaceTracking.__call__`` on a single (H,W,3) frame indexes ``frames[0]`` (:459-461); here the
resolved ``max_age`` / ``min_hits`` are used and a single frame is treated as a batch of one.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

_F = np.eye(7)
_F[0, 4] = _F[1, 5] = _F[2, 6] = 1.0                       # constant velocity of (cx, cy, area)
_H = np.eye(4, 7)
_R = np.diag([1.0, 1.0, 10.0, 10.0])
_Q = np.diag([1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0001])
_P0 = np.diag([10.0] * 4 + [10000.0] * 3)                  # unobserved initial velocities
_I7 = np.eye(7)


def _measure(boxes):
    """(K,4) corner boxes -> (K,4) measurements (cx, cy, area, ratio)."""
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    return np.stack([b[:, 0] + w / 2.0, b[:, 1] + h / 2.0, w * h, w / h], axis=1)


def _corners(x):
    """(T,7) states -> (T,4) corner boxes."""
    w = np.sqrt(x[:, 2] * x[:, 3])
    h = x[:, 2] / w
    return np.stack([x[:, 0] - w / 2.0, x[:, 1] - h / 2.0, x[:, 0] + w / 2.0, x[:, 1] + h / 2.0], axis=1)


def _iou_matrix(faces, tracks):
    """(F,4) x (T,4) -> (F,T) IoU, stored as float32 like the reference's matrix (:227-232)."""
    f, t = faces[:, None, :], tracks[None, :, :]
    w = np.maximum(0.0, np.minimum(f[..., 2], t[..., 2]) - np.maximum(f[..., 0], t[..., 0]))
    h = np.maximum(0.0, np.minimum(f[..., 3], t[..., 3]) - np.maximum(f[..., 1], t[..., 1]))
    inter = w * h
    area_f = (f[..., 2] - f[..., 0]) * (f[..., 3] - f[..., 1])
    area_t = (t[..., 2] - t[..., 0]) * (t[..., 3] - t[..., 1])
    with np.errstate(divide='ignore', invalid='ignore'):
        return (inter / (area_f + area_t - inter)).astype(np.float32)


class Sort:
    """Appearance-agnostic multi-face tracker (https://arxiv.org/abs/1602.00763): attaches an
    identity to every detection passed to it, or ``None`` while no confirmed track exists.
    Observations are returned as they are (no smoothing, no interpolation)."""

    #: ids are numbered across all instances of a process, like the reference's
    #: ``KalmanTracker.count`` (:113)
    next_id = 0

    def __init__(self, max_age=1, min_hits=3, return_unmatched=False, iou_threshold=0.3):
        self.max_age = max_age
        self.min_hits = min_hits
        self.return_unmatched = return_unmatched
        self.iou_threshold = iou_threshold
        self.frame_count = 0
        self.x = np.zeros((0, 7))            # states
        self.P = np.zeros((0, 7, 7))         # covariances
        self.hits = np.zeros(0, dtype=np.int64)
        self.since_update = np.zeros(0, dtype=np.int64)
        self.ids = np.zeros(0, dtype=np.int64)

    def __len__(self):
        return len(self.ids)

    def _keep(self, mask):
        self.x, self.P = self.x[mask], self.P[mask]
        self.hits, self.since_update, self.ids = self.hits[mask], self.since_update[mask], self.ids[mask]

    def _predict(self):
        """Advance every track one frame; returns the predicted (T,4) boxes."""
        shrink = self.x[:, 6] + self.x[:, 2] <= 0        # a box about to get a negative area:
        self.x[shrink, 6] = 0.0                          # stop its growth (:183-185)
        self.x = self.x @ _F.T
        self.P = _F @ self.P @ _F.T + _Q
        self.since_update += 1
        with np.errstate(invalid='ignore', divide='ignore'):
            return _corners(self.x)

    def _correct(self, which, boxes):
        """Kalman update of tracks `which` with their matched corner boxes."""
        z = _measure(boxes)                               # (K,4)
        x, P = self.x[which], self.P[which]
        y = z - x @ _H.T
        PHT = P @ _H.T                                    # (K,7,4)
        S = _H @ PHT + _R
        K = PHT @ np.linalg.inv(S)
        self.x[which] = x + (K @ y[:, :, None])[:, :, 0]
        I_KH = _I7 - K @ _H
        self.P[which] = I_KH @ P @ I_KH.transpose(0, 2, 1) + K @ _R @ K.transpose(0, 2, 1)
        self.hits[which] += 1
        self.since_update[which] = 0

    def update(self, faces):
        """Call on EVERY frame (also without detections).  faces: list of dicts as returned by
        ``Detection``; returns the same dicts with a ``track`` field (int or None), matched
        tracks first (in track order), then the new ones; unconfirmed ones are filtered unless
        ``return_unmatched``."""
        self.frame_count += 1
        boxes = self._predict()
        alive = ~np.isnan(boxes).any(axis=1)             # a track may extrapolate to infinity
        if not alive.all():
            self._keep(alive)
            boxes = boxes[alive]

        n_faces, n_tracks = len(faces), len(boxes)
        face_boxes = np.array([f['bbox'] for f in faces], dtype=np.float64).reshape(-1, 4)
        face_of_track = np.full(n_tracks, -1)
        if n_tracks and n_faces:
            iou = _iou_matrix(face_boxes, boxes)
            rows, cols = linear_sum_assignment(-iou)
            good = ~(iou[rows, cols] < self.iou_threshold)
            face_of_track[cols[good]] = rows[good]
            # new identities: never-assigned faces in index order, then the rejected matches
            # in assignment order (:237-259)
            taken = np.zeros(n_faces, dtype=bool)
            taken[rows] = True
            births = list(np.flatnonzero(~taken)) + list(rows[~good])
        else:
            births = list(range(n_faces))

        out = []
        matched = np.flatnonzero(face_of_track >= 0)
        if len(matched):
            self._correct(matched, face_boxes[face_of_track[matched]])
            confirmed = (self.hits[matched] >= self.min_hits) | (self.frame_count <= self.min_hits)
            for t, ok in zip(matched, confirmed):
                out.append({'track': int(self.ids[t]) if ok else None, **faces[face_of_track[t]]})

        if births:
            k = len(births)
            new_ids = np.arange(Sort.next_id, Sort.next_id + k)
            Sort.next_id += k
            x0 = np.zeros((k, 7))
            x0[:, :4] = _measure(face_boxes[births])
            self.x = np.concatenate([self.x, x0])
            self.P = np.concatenate([self.P, np.repeat(_P0[None], k, axis=0)])
            self.hits = np.concatenate([self.hits, np.zeros(k, dtype=np.int64)])
            self.since_update = np.concatenate([self.since_update, np.zeros(k, dtype=np.int64)])
            self.ids = np.concatenate([self.ids, new_ids])
            for f, i in zip(births, new_ids):
                out.append({'track': int(i) if self.min_hits == 0 else None, **faces[f]})

        if not self.return_unmatched:
            out = [f for f in out if f['track'] is not None]
        self._keep(self.since_update <= self.max_age)
        return out


class FaceTracking:
    """Used like a ``Detection`` object on same-size batches of frames; every face dict gets a
    ``track`` field.  Holds the detector and the ``Sort`` state (reference :414-468)."""

    def __init__(self, detector=None, tracker=None):
        self.detector = detector
        self.tracker = tracker

    def __call__(self, frames):
        single = not isinstance(frames, list) and np.ndim(frames) == 3
        if single:
            frames = np.asarray(frames)[None]
        per_frame = [self.tracker.update(faces) for faces in self.detector(frames)]
        return per_frame[0] if single else per_frame


def face_tracking(*, video=None, max_age=None, min_hits=None, detector=None, return_unmatched=False):
    """Factory of a ``FaceTracking`` instance (reference :471-552).  ``video`` (anything with a
    ``framerate``) derives ``max_age`` = one second and ``min_hits`` = a fifth of a second;
    explicit values take precedence; the defaults assume 30 fps."""
    from terran_b200.face.detection import Detection, face_detection
    max_age_, min_hits_ = 30, 6
    if video is not None:
        max_age_, min_hits_ = video.framerate, video.framerate // 5
    max_age = max_age_ if max_age is None else max_age
    min_hits = min_hits_ if min_hits is None else min_hits
    if detector is None:
        detector = face_detection
    elif not isinstance(detector, Detection):
        raise ValueError('`detector` must be an instance of `terran.face.Detection`.')
    return FaceTracking(detector=detector,
                        tracker=Sort(max_age=max_age, min_hits=min_hits, return_unmatched=return_unmatched))
