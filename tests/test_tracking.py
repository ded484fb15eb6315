"""CPU: SORT face tracking.  The oracle restatement is pinned to the
reference's own ``Sort`` by ``tests/golden/sort_tracking.npz`` (made by
``oracle/make_golden_track.py`` from ``/root/reference/terran/tracking/face.py``); the product's
vectorised tracker must reproduce the oracle's identities, filtering and output order."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import track
from terran_b200.tracking import FaceTracking, Sort, face_tracking

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'sort_tracking.npz')


def encode(per_frame, id0=0):
    rows = []
    for t, faces in enumerate(per_frame):
        for f in faces:
            rows.append([t, f['person'], -1 if f['track'] is None else f['track'] - id0, *f['bbox']])
    return np.array(rows, dtype=np.int64).reshape(-1, 7)


def golden_cases():
    g = np.load(GOLDEN)
    n = len([k for k in g.files if k.startswith('case')])
    return [(g[f'cfg{i}'], g[f'case{i}']) for i in range(n)]


@pytest.mark.parametrize('i', range(6))
def test_oracle_matches_reference_sort(i):
    cfg, want = golden_cases()[i]
    seed, frames, people, max_age, min_hits, ru = (int(v) for v in cfg)
    seq = track.synthetic_sequence(seed, frames, people)
    o = track.SortOracle(max_age, min_hits, bool(ru))
    np.testing.assert_array_equal(encode([o.update(f) for f in seq]), want)


@pytest.mark.parametrize('i', range(6))
def test_product_matches_reference_sort(i):
    cfg, want = golden_cases()[i]
    seed, frames, people, max_age, min_hits, ru = (int(v) for v in cfg)
    seq = track.synthetic_sequence(seed, frames, people)
    id0 = Sort.next_id
    s = Sort(max_age=max_age, min_hits=min_hits, return_unmatched=bool(ru))
    np.testing.assert_array_equal(encode([s.update(f) for f in seq], id0), want)


@pytest.mark.parametrize('seed', range(20))
def test_product_matches_oracle_on_random_sequences(seed):
    rng = np.random.default_rng(100 + seed)
    people, frames = int(rng.integers(1, 12)), int(rng.integers(5, 60))
    max_age, min_hits = int(rng.integers(1, 8)), int(rng.integers(0, 5))
    ru = bool(rng.integers(0, 2))
    seq = track.synthetic_sequence(1000 + seed, frames, people, jitter=float(rng.uniform(0, 8)),
                                   drop=float(rng.uniform(0, 0.5)))
    seq[int(rng.integers(0, frames))] = []                 # a frame without detections
    o = track.SortOracle(max_age, min_hits, ru)
    id0 = Sort.next_id
    s = Sort(max_age=max_age, min_hits=min_hits, return_unmatched=ru)
    for t, faces in enumerate(seq):
        a, b = o.update(faces), s.update(faces)
        np.testing.assert_array_equal(encode([a]), encode([b], id0), err_msg=f'frame {t}')
        assert len(s) == len(o.tracks)
    # the filter state itself agrees to rounding
    if len(s):
        want = np.stack([t.kf.x[:, 0] for t in o.tracks])
        np.testing.assert_allclose(s.x, want, rtol=1e-9, atol=1e-9)


def test_tracks_follow_people_through_dropouts():
    seq = track.synthetic_sequence(7, frames=50, people=3, jitter=1.0, drop=0.1)
    s = Sort(max_age=5, min_hits=2)
    owner = {}
    for faces in seq:
        for f in s.update(faces):
            owner.setdefault(f['track'], set()).add(f['person'])
    assert all(len(p) == 1 for p in owner.values())        # an identity never jumps between people
    assert len(owner) <= 5                                 # and people are not re-numbered at every dropout


def test_output_dicts_keep_detection_fields_and_are_not_mutated():
    faces = [{'bbox': np.array([10, 10, 60, 70], np.int32), 'landmarks': np.zeros((5, 2), np.int32),
              'score': np.float32(0.9)}]
    s = Sort(max_age=1, min_hits=0)
    out = s.update(faces)
    assert list(out[0]) == ['track', 'bbox', 'landmarks', 'score'] and isinstance(out[0]['track'], int)
    assert 'track' not in faces[0]
    assert s.update([]) == [] and len(s) == 1              # kept for max_age frames
    assert s.update([]) == [] and len(s) == 0


class _FakeDetection:
    def __call__(self, frames):
        return [[{'bbox': np.array([5 + i, 5, 45 + i, 55], np.int32), 'score': np.float32(1)}]
                for i in range(len(frames))]


def test_face_tracking_wrapper_and_factory(monkeypatch):
    ft = FaceTracking(detector=_FakeDetection(), tracker=Sort(max_age=3, min_hits=0))
    frames = np.zeros((4, 64, 64, 3), np.uint8)
    per_frame = ft(frames)
    assert len(per_frame) == 4 and len({f[0]['track'] for f in per_frame}) == 1
    single = ft(frames[0])                                 # one (H,W,3) frame -> one list of faces
    assert isinstance(single, list) and single[0]['track'] == per_frame[0][0]['track']
    with pytest.raises(ValueError, match='must be an instance'):
        face_tracking(detector=object())

    class Video:
        framerate = 25
    import terran_b200.face.detection as det
    monkeypatch.setattr(det, 'face_detection', _FakeDetection())
    t = face_tracking(video=Video())
    assert (t.tracker.max_age, t.tracker.min_hits) == (25, 5)
    t = face_tracking(video=Video(), max_age=7)
    assert (t.tracker.max_age, t.tracker.min_hits) == (7, 5)
    t = face_tracking()                                    # upstream crashes here (video is None)
    assert (t.tracker.max_age, t.tracker.min_hits) == (30, 6)


TRACK_WORKER = """
import sys
sys.path.insert(0, {root!r})
import numpy as np
from oracle import track
from terran_b200 import parallel
from terran_b200.tracking import Sort

rank, world, _ = parallel.init_from_env('gloo')
seq = track.synthetic_sequence(11, frames=33, people=5)
for faces in seq:
    for f in faces:
        f['landmarks'] = np.tile(f['bbox'][:2], (5, 1)).astype(np.int32)
lo, hi = parallel.shard_range(len(seq), rank, world)
tracked = parallel.track_sharded(Sort(max_age=4, min_hits=2), seq[lo:hi])
if rank == 0:
    id0 = Sort.next_id
    s = Sort(max_age=4, min_hits=2)
    want = [s.update([dict((k, v) for k, v in f.items() if k != 'person') for f in faces]) for faces in seq]
    assert len(tracked) == len(want) == 33
    for a, b in zip(tracked, want):
        assert len(a) == len(b)
        for fa, fb in zip(a, b):
            # two trackers of one process: ids differ by the offset between the two runs
            assert (fa['track'] is None) == (fb['track'] is None)
            assert np.array_equal(fa['bbox'], fb['bbox']) and np.array_equal(fa['landmarks'], fb['landmarks'])
            assert fa['bbox'].dtype == np.int32 and fa['score'] == fb['score']
    ids_a = [f['track'] for fr in tracked for f in fr if f['track'] is not None]
    ids_b = [f['track'] for fr in want for f in fr if f['track'] is not None]
    assert len(ids_a) > 20 and [i - min(ids_a) for i in ids_a] == [i - min(ids_b) for i in ids_b]
    print('TRACK_OK')
else:
    assert tracked is None
"""


def test_two_rank_gloo_tracking_consumes_gathered_frames_in_order(tmp_path):
    """Frames shard across ranks, tracking runs on rank 0 over the gathered detections in frame
    order and equals single-process tracking of the whole sequence."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'worker.py'
    script.write_text(TRACK_WORKER.format(root=root))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    r = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29733', str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'TRACK_OK' in r.stdout
