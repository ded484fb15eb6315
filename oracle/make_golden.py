"""ORACLE tooling: run the UNMODIFIED reference (imported from /root/reference,
build container only) on seeded synthetic inputs and weights, check the
restatements in ``oracle/`` against it, and write the golden fixtures under
``tests/golden/``.

    python oracle/make_golden.py            # regenerate fixtures + pin report

The reference cannot travel to the GPU box, so the fixtures are what pins the
oracle there.  Shims: stub ``ffmpeg`` / ``skimage``
modules so ``import terran`` works; a ``.contiguous()`` wrapper around the
RetinaFace module (the reference's channels-last ``.view`` crashes on this
torch); synthetic checkpoints under a scratch ``TERRAN_HOME``; CPU only.
"""
import os
import sys
import tempfile
import types

os.environ['CUDA_VISIBLE_DEVICES'] = ''
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get('TERRAN_REFERENCE', '/root/reference')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from terran_b200 import synth  # noqa: E402
from oracle import nets, detect, pose  # noqa: E402


def import_reference():
    for name in ('ffmpeg', 'skimage', 'skimage.transform'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['skimage.transform'].SimilarityTransform = type(
        'SimilarityTransform', (), {})
    home = tempfile.mkdtemp(prefix='terran_home_')
    os.makedirs(os.path.join(home, 'checkpoints'))
    os.environ['TERRAN_HOME'] = home
    torch.save(synth.retinaface_state_dict(), os.path.join(home, 'checkpoints', 'b5d77fff.pth'))
    torch.save(synth.arcface_state_dict(), os.path.join(home, 'checkpoints', 'd206e4b0.pth'))
    torch.save(synth.openpose_state_dict(), os.path.join(home, 'checkpoints', '11a769ad.pth'))
    sys.path.insert(0, REFERENCE)
    import terran  # noqa: F401
    return terran


class Contiguous(torch.nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        return self.inner(x.contiguous())


def report(name, a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = float(np.abs(a - b).max()) if a.size else 0.0
    print(f'  pin {name:42s} max|diff| = {d:.3e}  (range {np.abs(a).max() if a.size else 0:.3g})')
    return d


def golden_retinaface():
    from terran.face.detection import Detection
    from terran.face.detection.retinaface import RetinaFace
    sd = synth.retinaface_state_dict()

    wrapper = RetinaFace(device=torch.device('cpu'))
    wrapper.model = Contiguous(wrapper.model)

    # (a) network forward on a small odd-sized batch.
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.integers(0, 256, (2, 3, 75, 109)).astype(np.float32))
    with torch.no_grad():
        ref_heads = [t.numpy() for t in wrapper.model(x)]
    ora_heads = [t.numpy() for t in nets.retinaface_forward(sd, x)]
    for i, (a, b) in enumerate(zip(ref_heads, ora_heads)):
        assert report(f'retinaface head[{i}] {a.shape}', a, b) < 2e-5
    np.savez_compressed(os.path.join(GOLDEN, 'retinaface_forward.npz'),
                        x=x.numpy().astype(np.uint8), **{f'head{i}': h for i, h in enumerate(ref_heads)})

    # (b) RetinaFace.call (decode + threshold + sort + NMS) on u8 frames.
    images = rng.integers(0, 256, (3, 160, 232, 3), dtype=np.uint8)
    ref = wrapper.call(images)
    xin = torch.from_numpy(images.astype(np.float32)).permute(0, 3, 1, 2).flip(1)
    with torch.no_grad():
        heads = [t.numpy() for t in wrapper.model(xin)]
    ora = detect.model_call(heads, 160, 232)
    fx = {'images': images}
    for i, (r, o) in enumerate(zip(ref, ora)):
        print(f'  image {i}: reference {len(r)} faces, oracle {len(o)} faces')
        assert len(r) == len(o) and len(r) > 0
        for key in ('bbox', 'landmarks', 'score'):
            a = np.stack([f[key] for f in r])
            b = np.stack([f[key] for f in o])
            if key == 'score':
                assert np.array_equal(a, b), (i, key)   # same survivors, same order
            else:                                       # exp() differs by <= 1 ulp
                assert np.allclose(a, b, rtol=0, atol=2e-4), (i, key, np.abs(a - b).max())
            fx[f'{key}{i}'] = a
    for i, h in enumerate(heads):
        fx[f'head{i}'] = h
    np.savez_compressed(os.path.join(GOLDEN, 'retinaface_call.npz'), **fx)

    # (c) Detection.__call__ end to end on one 640x640 image (BASELINE config 1).
    det = Detection(device=torch.device('cpu'), lazy=True)
    det.model = wrapper
    img = np.random.default_rng(0).integers(0, 256, (640, 640, 3), dtype=np.uint8)
    faces = det(img)
    print(f'  Detection()(640x640): {len(faces)} faces')
    np.savez_compressed(
        os.path.join(GOLDEN, 'retinaface_detection_640.npz'),
        bbox=np.stack([f['bbox'] for f in faces]),
        landmarks=np.stack([f['landmarks'] for f in faces]),
        score=np.stack([f['score'] for f in faces]))


def golden_arcface():
    from terran.face.recognition import Recognition
    from terran.face.recognition.arcface import ArcFace
    sd = synth.arcface_state_dict()
    wrapper = ArcFace(device=torch.device('cpu'))
    rng = np.random.default_rng(21)
    crops = rng.integers(0, 256, (3, 112, 112, 3), dtype=np.uint8)
    rec = Recognition(device=torch.device('cpu'), lazy=True)
    rec.model = wrapper
    ref = rec(list(crops))
    x = torch.from_numpy(crops.transpose(0, 3, 1, 2)[:, ::-1].astype(np.float32).copy())
    with torch.no_grad():
        raw_ref = wrapper.model(x).numpy()
    raw = nets.arcface_forward(sd, x).numpy()
    assert report('arcface raw embedding', raw_ref, raw) < 1e-4
    norm = np.sqrt((raw.astype(np.float32) ** 2).sum(1, keepdims=True))
    norm[norm == 0] = 1
    assert report('arcface normalised', ref, raw / norm) < 1e-6
    print(f'  embedding raw range {np.abs(raw_ref).max():.3f}')
    np.savez_compressed(os.path.join(GOLDEN, 'arcface_embed.npz'), crops=crops,
                        raw=raw_ref, normalised=np.asarray(ref, np.float32))


def golden_openpose():
    from terran.pose import Estimation
    from terran.pose.openpose import OpenPose
    sd = synth.openpose_state_dict()
    wrapper = OpenPose(device=torch.device('cpu'))

    rng = np.random.default_rng(31)
    x = torch.from_numpy((rng.integers(0, 256, (2, 3, 56, 72)) / 255.0 - 0.5).astype(np.float32))
    with torch.no_grad():
        paf_ref, heat_ref = (t.numpy() for t in wrapper.model(x))
    paf, heat = (t.numpy() for t in nets.openpose_forward(sd, x))
    assert report('openpose paf', paf_ref, paf) < 1e-5
    assert report('openpose heat', heat_ref, heat) < 1e-5
    np.savez_compressed(os.path.join(GOLDEN, 'openpose_forward.npz'), x=x.numpy(),
                        paf=paf_ref, heat=heat_ref)

    # bicubic x8 restatement vs torch.
    t = torch.from_numpy(rng.random((1, 5, 23, 40)).astype(np.float32))
    up = torch.nn.functional.interpolate(t, scale_factor=8, mode='bicubic',
                                         align_corners=False)[0].numpy()
    assert report('bicubic x8', up, pose.bicubic_up8(t[0].numpy())) < 1e-6

    # Parse: synthetic maps injected through a stub module; the reference's
    # OpenPose.call runs its decode unchanged.
    class Stub(torch.nn.Module):
        def forward(self, _x):
            return self.out

    stub = Stub()
    wrapper.model = stub
    est = Estimation(device=torch.device('cpu'), lazy=True)
    est.model = wrapper
    frames = np.zeros((1, 720, 1280, 3), np.uint8)
    scale = 184 / 720
    fx = {}
    n_humans = []
    for scene in range(24):
        paf, heat = pose.synthetic_scene(1000 + scene)
        stub.out = (torch.from_numpy(paf)[None], torch.from_numpy(heat)[None])
        ref = est(frames)[0]
        ora = pose.parse_frame(paf, heat, scale)
        assert len(ref) == len(ora), (scene, len(ref), len(ora))
        for r, o in zip(ref, ora):
            assert np.array_equal(r['keypoints'], o['keypoints']), scene
            assert abs(r['score'] - o['score']) < 1e-5, scene
        n_humans.append(len(ref))
        if scene < 4:     # full maps for a few scenes; the rest regenerate from the seed
            fx[f'paf{scene}'], fx[f'heat{scene}'] = paf.astype(np.float16), heat.astype(np.float16)
        fx[f'mapsum{scene}'] = np.array([paf.astype(np.float64).sum(), heat.astype(np.float64).sum()])
        fx[f'kp{scene}'] = (np.stack([r['keypoints'] for r in ref]) if ref
                            else np.zeros((0, 18, 3), np.int32))
        fx[f'score{scene}'] = np.array([r['score'] for r in ref], np.float64)
    print('  parse scenes humans:', n_humans)
    fx['seeds'] = np.arange(1000, 1024)
    np.savez_compressed(os.path.join(GOLDEN, 'openpose_parse.npz'), **fx)


def golden_baseline():
    """Fixtures at the BASELINE.json batch sizes, produced by the unmodified reference:
    C3 = Recognition on 256 crops of 112x112 (the bench's crops), C4/C5 = Estimation on 720p
    noise frames with the peak-calibrated OpenPose checkpoint (humans in every frame), C2 =
    Detection on four 1080p noise frames."""
    from terran.face.detection import Detection
    from terran.face.detection.retinaface import RetinaFace
    from terran.face.recognition import Recognition
    from terran.face.recognition.arcface import ArcFace
    from terran.pose import Estimation
    from terran.pose.openpose import OpenPose

    # -- C3: 256 crops
    crops = np.random.default_rng(2).integers(0, 256, (256, 112, 112, 3), dtype=np.uint8)
    rec = Recognition(device=torch.device('cpu'), lazy=True)
    rec.model = ArcFace(device=torch.device('cpu'))
    ref = np.asarray(rec(list(crops)), np.float32)
    sd = synth.arcface_state_dict()
    x = torch.from_numpy(crops[:4].transpose(0, 3, 1, 2)[:, ::-1].astype(np.float32).copy())
    raw = nets.arcface_forward(sd, x).numpy()
    assert report('arcface b256 (first 4) normalised', ref[:4],
                  raw / np.linalg.norm(raw, axis=1, keepdims=True)) < 1e-6
    np.savez_compressed(os.path.join(GOLDEN, 'arcface_embed_b256.npz'), normalised=ref,
                        crops_seed=np.array(2))

    # -- C4: Estimation with humans (checkpoint swapped for the calibrated one)
    sdp = synth.openpose_state_dict(peaks=True)
    wrapper = OpenPose(device=torch.device('cpu'))
    wrapper.model.load_state_dict(sdp)
    est = Estimation(device=torch.device('cpu'), lazy=True)
    est.model = wrapper
    frames = np.random.default_rng(1).integers(0, 256, (2, 720, 1280, 3), dtype=np.uint8)
    ref = est(frames)
    import cv2
    s = 184 / 720
    small = np.stack([cv2.resize(f, (int(1280 * s), int(720 * s)), interpolation=cv2.INTER_LINEAR)
                      for f in frames])
    xin = torch.from_numpy(small.transpose(0, 3, 1, 2).astype(np.float32) / 255.0 - 0.5)
    paf, heat = nets.openpose_forward(sdp, xin)
    ora = pose.parse(paf.numpy(), heat.numpy(), s)
    fx = {'frames_seed': np.array(1)}
    for n, (r, o) in enumerate(zip(ref, ora)):
        print(f'  Estimation 720p frame {n}: reference {len(r)} humans, oracle {len(o)}')
        assert len(r) == len(o) and len(r) > 0
        for a, b in zip(r, o):
            assert np.array_equal(a['keypoints'], b['keypoints'])
            assert abs(a['score'] - b['score']) < 1e-5
        fx[f'kp{n}'] = np.stack([a['keypoints'] for a in r])
        fx[f'score{n}'] = np.array([a['score'] for a in r], np.float64)
    np.savez_compressed(os.path.join(GOLDEN, 'openpose_estimation_720p.npz'), **fx)

    # -- C2: Detection on 1080p frames
    wrapper = RetinaFace(device=torch.device('cpu'))
    wrapper.model = Contiguous(wrapper.model)
    det = Detection(device=torch.device('cpu'), lazy=True)
    det.model = wrapper
    frames = np.random.default_rng(0).integers(0, 256, (4, 1080, 1920, 3), dtype=np.uint8)
    ref = det(frames)
    fx = {'frames_seed': np.array(0)}
    for n, faces in enumerate(ref):
        print(f'  Detection 1080p frame {n}: {len(faces)} faces')
        fx[f'bbox{n}'] = np.stack([f['bbox'] for f in faces])
        fx[f'landmarks{n}'] = np.stack([f['landmarks'] for f in faces])
        fx[f'score{n}'] = np.stack([f['score'] for f in faces])
    np.savez_compressed(os.path.join(GOLDEN, 'retinaface_detection_1080p.npz'), **fx)


def golden_letterbox():
    """Faces WITHOUT landmarks: the reference's own ``preprocess_face_no_landmarks``
    (arcface/wrapper.py:75-99) on seeded images of many sizes; checks ``oracle/letterbox.py``."""
    from terran.face.recognition.arcface.wrapper import preprocess_face_no_landmarks
    from oracle import letterbox as lb
    rng = np.random.default_rng(21)
    shapes = [(112, 112), (57, 41), (300, 181), (181, 300), (9, 640), (640, 9), (3, 5),
              (111, 113), (1, 1), (217, 212), (270, 480)]
    fx, worst = {}, 0.0
    for i, (h, w) in enumerate(shapes):
        image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i % 3 == 0 or h * w > 100000:        # smooth content: long filter windows over gradients (and small fixtures)
            yy, xx = np.mgrid[0:h, 0:w]
            image = np.stack([xx * 255 // max(w - 1, 1), yy * 255 // max(h - 1, 1), (xx + yy) % 256],
                             -1).astype(np.uint8)
        want = preprocess_face_no_landmarks(image, 112)
        worst = max(worst, report(f'letterbox {h}x{w}', lb.letterbox(image), want))
        fx[f'image_{i}'], fx[f'crop_{i}'] = image, want
    assert worst == 0.0, 'oracle/letterbox.py differs from the reference'
    fx['n'] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(GOLDEN, 'letterbox.npz'), **fx)


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    import_reference()
    which = sys.argv[1:] or ['retinaface', 'arcface', 'openpose', 'baseline', 'letterbox']
    for name in which:
        print(f'[{name}]')
        globals()['golden_' + name]()
    print('golden fixtures written to', GOLDEN)
