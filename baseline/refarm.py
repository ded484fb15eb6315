"""The reference arm: the UNMODIFIED reference (``terran`` from ``baseline/_ref``, installed
with ``pip install --no-deps --target baseline/_ref /root/reference``; or ``/root/reference``
itself in the build container) driven through its own public wrappers on the CPU.

Three shims, none of which touches reference code:
  1. stub modules for the absent ``ffmpeg`` (video I/O) and ``skimage`` (5-point alignment)
     dependencies so that ``import terran`` succeeds — neither is on the detect/pose path;
  2. a ``.contiguous()`` wrapper around the RetinaFace ``nn.Module`` (the reference's
     channels-last ``.view`` raises on torch >= 2.x);
  3. ``TERRAN_HOME`` pointing at a scratch directory holding the synthetic checkpoints under
     the reference's own checkpoint ids, CPU only (``CUDA_VISIBLE_DEVICES=''`` must be set by
     the caller BEFORE torch is imported: the reference sends anchors and pose inputs to its
     ``default_device`` regardless of the ``device`` argument).

Used by ``bench.py --impl reference`` / ``cpu_baseline`` and by the boundary tests; nothing in
``terran_b200/`` imports this module.
"""
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference')

CHECKPOINT_IDS = {'retinaface': 'b5d77fff', 'arcface': 'd206e4b0', 'openpose': '11a769ad'}


def reference_path():
    for path in CANDIDATES:
        if os.path.isdir(os.path.join(path, 'terran')):
            return path
    return None


def import_reference(state_dicts=None):
    """Import the reference package with the shims in place.  ``state_dicts`` maps
    'retinaface' / 'arcface' / 'openpose' to the checkpoints to serve (default: the seeded
    synthetic ones).  Returns the ``terran`` module."""
    import torch
    path = reference_path()
    if path is None:
        raise ImportError('the reference is not installed (baseline/_ref) nor at /root/reference')
    for name in ('ffmpeg', 'skimage', 'skimage.transform'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules['skimage.transform'], 'SimilarityTransform'):
        sys.modules['skimage.transform'].SimilarityTransform = type('SimilarityTransform', (), {})
    if 'terran' not in sys.modules:
        from terran_b200 import synth
        home = tempfile.mkdtemp(prefix='terran_ref_home_')
        os.makedirs(os.path.join(home, 'checkpoints'))
        os.environ['TERRAN_HOME'] = home
        defaults = {'retinaface': synth.retinaface_state_dict, 'arcface': synth.arcface_state_dict,
                    'openpose': synth.openpose_state_dict}
        for name, cid in CHECKPOINT_IDS.items():
            sd = (state_dicts or {}).get(name)
            torch.save(sd if sd is not None else defaults[name](),
                       os.path.join(home, 'checkpoints', cid + '.pth'))
        sys.path.insert(0, path)
    import terran
    return terran


class Contiguous:
    """Shim 2 as a callable module wrapper (built lazily so torch is imported by the caller)."""

    def __new__(cls, inner):
        import torch

        class _Contiguous(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.inner = m

            def forward(self, x):
                return self.inner(x.contiguous())

        return _Contiguous(inner)


def reference_detection(state_dict=None):
    """The reference's ``Detection`` on the CPU, ready to call."""
    import torch
    import_reference()
    from terran.face.detection import Detection
    from terran.face.detection.retinaface import RetinaFace
    model = RetinaFace(device=torch.device('cpu'))
    if state_dict is not None:
        model.model.load_state_dict(state_dict)
    model.model = Contiguous(model.model)
    det = Detection(device=torch.device('cpu'), lazy=True)
    det.model = model
    return det


def reference_estimation(state_dict=None):
    """The reference's ``Estimation`` on the CPU, ready to call."""
    import torch
    import_reference()
    from terran.pose import Estimation
    from terran.pose.openpose import OpenPose
    model = OpenPose(device=torch.device('cpu'))
    if state_dict is not None:
        model.model.load_state_dict(state_dict)
    est = Estimation(device=torch.device('cpu'), lazy=True)
    est.model = model
    return est
