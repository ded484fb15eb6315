"""The drop-in boundary exercised through the REAL reference:
``terran_b200.checkpoint.register_with_reference()`` appends the B200 classes to the reference's
own ``CHECKPOINTS`` registry under the alias 'b200', after which the unmodified
``terran.face.Detection(checkpoint='b200')`` etc. resolve to — and, on a GPU, run — them
(reference: terran/checkpoint.py:213-245, face/detection/__init__.py:220,276-278).

Runs in a child process (the reference creates ~/.terran-style state at import and other tests
install a fake ``terran`` module); skipped where neither baseline/_ref nor /root/reference exists.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refarm  # noqa: E402

needs_reference = pytest.mark.skipif(refarm.reference_path() is None,
                                     reason='reference not installed (baseline/_ref)')

B200 = ['terran_b200.face.detection.retinaface.wrapper.RetinaFace',
        'terran_b200.face.recognition.arcface.wrapper.ArcFace',
        'terran_b200.pose.openpose.wrapper.OpenPose']


def run_child(mode):
    env = {k: v for k, v in os.environ.items() if k != 'TERRAN_HOME'}
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_plugin_script.py'), mode],
                       capture_output=True, text=True, timeout=900, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1][7:])


def check_registry(out):
    assert out['n_b200_entries'] == 3
    assert out['classes'] == B200
    assert all(c.startswith('terran.') for c in out['defaults']), out['defaults']   # defaults untouched
    assert out['repr'] == '<Detection(RetinaFace)>'
    assert out['bad_alias'] == 'Checkpoint not found.'


@needs_reference
def test_reference_resolves_alias_b200_to_the_b200_classes():
    check_registry(run_child('cpu'))


@needs_reference
@pytest.mark.gpu
def test_reference_wrappers_run_the_b200_classes():
    out = run_child('gpu')
    check_registry(out)
    assert out['faces_equal'] and sum(out['faces']) > 10, out
    assert out['humans_equal'] and sum(out['humans']) > 3, out
    assert out['single_equal'] and len(out['list_lens']) == 2
    assert out['emb_shape'] == [4, 512] and out['emb_norm_ok']
