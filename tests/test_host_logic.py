"""CPU: host-side logic — registry, batching, weight packing, the C-ABI library
(loads, exports every declared symbol), and the multi-process sharding helpers
over gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.conftest import ROOT


# ----------------------------------------------------------------- C ABI / build

def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, 'include', 'terran_b200.h')).read()
    declared = set(re.findall(r'\b(tr_[a-z0-9_]+)\s*\(', header))
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tr_version() == 100


def test_op_desc_layout_matches_header(native):
    """ctypes struct layout == the C struct (field order and sizes)."""
    header = open(os.path.join(ROOT, 'include', 'terran_b200.h')).read()
    body = re.search(r'typedef struct tr_op_desc \{(.*?)\} tr_op_desc;', header, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        fields += [(n.strip(), ctype) for n in names.split(',')]
    want = [(n.rstrip('_'), {'int32_t': C.c_int32, 'int64_t': C.c_int64, 'float': C.c_float}[t])
            for n, t in fields]
    got = [(n.rstrip('_'), t) for n, t in native.OpDesc._fields_]
    assert got == want
    assert C.sizeof(native.OpDesc) == 22 * 4 + 6 * 8 + 2 * 4 + 4 * 8 + 2 * 4 + 8


def test_no_gpu_calls_fail_loudly(native):
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(native.NativeError):
        native.check(native.lib().tr_init(0))
    from terran_b200.face.detection.retinaface import RetinaFace
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        RetinaFace(device=torch.device('cpu'), state_dict={})


def test_bicubic_table_matches_oracle(native):
    from oracle import pose
    tab = (C.c_float * 32)()
    native.lib().tr_bicubic_table(tab)
    np.testing.assert_array_equal(np.array(tab, np.float32).reshape(8, 4), pose.bicubic_table())


def test_sources_build_for_sm100a():
    from terran_b200 import build
    src = open(os.path.join(ROOT, 'terran_b200', 'build.py')).read()
    assert 'arch=compute_100a,code=sm_100a' in src and '-lineinfo' in src
    assert os.path.exists(build.build())


# ------------------------------------------------------------------- registry

def test_registry_resolves_classes(tmp_path, monkeypatch):
    monkeypatch.setenv('TERRAN_HOME', str(tmp_path))
    from terran_b200 import checkpoint as ck
    from terran_b200.face.detection.retinaface import RetinaFace
    from terran_b200.pose.openpose import OpenPose
    assert ck.get_class_for_checkpoint('face-detection', None) is RetinaFace
    assert ck.get_class_for_checkpoint('pose-estimation', 'gpu-realtime') is OpenPose
    with pytest.raises(ValueError, match='Checkpoint not found'):
        ck.get_class_for_checkpoint('face-detection', 'nope')
    with pytest.raises(ValueError):
        ck.get_checkpoint_path('terran_b200.pose.openpose.OpenPose')       # not on disk
    (tmp_path / 'checkpoints').mkdir(exist_ok=True)
    (tmp_path / 'checkpoints' / '11a769ad.pth').write_bytes(b'x')
    assert ck.get_checkpoint_path('terran_b200.pose.openpose.OpenPose').name == '11a769ad.pth'
    assert (tmp_path / 'checkpoints').exists()


def test_register_with_reference_registry(monkeypatch):
    """The plugin hook: entries appended to the reference's CHECKPOINTS list make
    its own lookup resolve alias 'b200' to these classes (terran/checkpoint.py:
    190-197, 244-245), without disturbing the defaults."""
    import types
    from terran_b200 import checkpoint as ck
    fake_pkg, fake = types.ModuleType('terran'), types.ModuleType('terran.checkpoint')
    fake.CHECKPOINTS = [{'id': 'b5d77fff', 'task': 'face-detection', 'alias': 'gpu-realtime',
                         'default': True, 'class': 'terran.face.detection.retinaface.RetinaFace'}]
    fake_pkg.checkpoint = fake
    monkeypatch.setitem(sys.modules, 'terran', fake_pkg)
    monkeypatch.setitem(sys.modules, 'terran.checkpoint', fake)
    ck.register_with_reference()
    ck.register_with_reference()          # idempotent
    added = [c for c in fake.CHECKPOINTS if c['alias'] == 'b200']
    assert [c['task'] for c in added] == ['face-detection', 'face-recognition', 'pose-estimation']
    assert all(not c['default'] for c in added)
    assert added[0]['id'] == 'b5d77fff' and added[0]['class'].startswith('terran_b200.')
    # the reference's own selection rule
    pick = [c for c in fake.CHECKPOINTS if c['task'] == 'face-detection' and c['alias'] == 'b200']
    assert len(pick) == 1
    default = [c for c in fake.CHECKPOINTS if c['task'] == 'face-detection' and c['default']]
    assert default[0]['class'] == 'terran.face.detection.retinaface.RetinaFace'


def test_public_surface():
    import terran_b200
    from terran_b200.pose import Keypoint
    for name in ('face_detection', 'extract_features', 'pose_estimation', 'Detection',
                 'Recognition', 'Estimation', 'default_device'):
        assert hasattr(terran_b200, name)
    assert repr(terran_b200.face_detection) == '<Detection(RetinaFace)>'
    assert repr(terran_b200.extract_features) == '<Recognition(ArcFace)>'
    assert repr(terran_b200.pose_estimation) == '<Estimation(OpenPose)>'
    assert Keypoint.NOSE.value == 0 and Keypoint.L_EAR.value == 17 and len(Keypoint) == 18


# ------------------------------------------------------------------- batching

def test_pad_merge_matches_reference_rules():
    from terran_b200.batching import PadMerge
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((10, 20), (13, 17), (8, 8))]
    batch, off = PadMerge().merge(imgs)
    assert batch.shape == (3, 13, 20, 3)
    # extra rows/cols: ceil before, floor after
    np.testing.assert_array_equal(off, [[0, 2], [2, 0], [6, 3]])
    for im, (left, top), b in zip(imgs, off, batch):
        np.testing.assert_array_equal(b[top:top + im.shape[0], left:left + im.shape[1]], im)
        assert b.sum() == im.sum()
    arr = np.zeros((2, 4, 4, 3), np.uint8)
    assert PadMerge('bogus').merge(arr)[1] is None          # arrays pass through unchecked
    with pytest.raises(NotImplementedError):
        PadMerge('crop').merge(imgs)
    with pytest.raises(ValueError):
        PadMerge('bogus').merge(imgs)

    faces = [[{'bbox': np.array([5, 6, 9, 12], np.float32),
               'landmarks': np.arange(10, dtype=np.float32).reshape(5, 2), 'score': np.float32(.9)}]] * 3
    out = PadMerge().unpad_faces(faces, off)
    np.testing.assert_array_equal(out[1][0]['bbox'], [3, 6, 7, 12])
    assert out[1][0]['bbox'].dtype == np.float32 and out[1][0]['landmarks'].dtype == np.float64
    poses = [[{'keypoints': np.array([[7, 9, 1], [0, 0, 0]] * 9, np.int32), 'score': 0.5}]] * 3
    kp = PadMerge().unpad_poses(poses, off)[2][0]['keypoints']
    np.testing.assert_array_equal(kp[0], [1, 6, 1])
    np.testing.assert_array_equal(kp[1], [0, 0, 0])          # absent joints stay zero


def test_round_faces_half_to_even():
    from terran_b200.batching import round_faces
    f = {'bbox': np.array([0.5, 1.5, 2.5, -0.5], np.float32) * np.float32(0.5),
         'landmarks': np.zeros((5, 2), np.float32), 'score': np.float32(1)}
    out = round_faces([[f]], 0.5)
    np.testing.assert_array_equal(out[0][0]['bbox'], [0, 2, 2, 0])
    assert out[0][0]['bbox'].dtype == np.int32


def test_resized_dims_follow_reference():
    from terran_b200.frames import resized_dims
    assert resized_dims(1080, 1920, 416)[:2] == (416, 739)
    assert resized_dims(720, 1280, 184)[:2] == (184, 327)
    assert resized_dims(640, 640, 416)[:2] == (416, 416)
    assert abs(resized_dims(1080, 1920, 416)[2] - 0.385185) < 1e-6


# ------------------------------------------------------------- weight packing

def test_bn_fold_is_exact_affine():
    from tests.reference_programs import bn_fold
    g = torch.Generator().manual_seed(0)
    sd = {'bn.weight': torch.rand(6, generator=g) + 0.5, 'bn.bias': torch.randn(6, generator=g),
          'bn.running_mean': torch.randn(6, generator=g), 'bn.running_var': torch.rand(6, generator=g) + 0.5}
    bias = torch.randn(6, generator=g)
    s, t = bn_fold(sd, 'bn', 2e-5, bias)
    y = torch.randn(4, 6, 3, 3, generator=g)
    want = torch.nn.functional.batch_norm(y + bias.view(1, -1, 1, 1), sd['bn.running_mean'],
                                          sd['bn.running_var'], sd['bn.weight'], sd['bn.bias'],
                                          False, 0.0, 2e-5)
    got = y * torch.from_numpy(s).view(1, -1, 1, 1) + torch.from_numpy(t).view(1, -1, 1, 1)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=1e-5)


def test_programs_cover_every_checkpoint_tensor():
    """Every conv / linear weight of the reference state_dicts lands in the
    packed blob exactly once (op counts per architecture)."""
    from terran_b200 import _native as nat, synth, weights
    P, roles = weights.retinaface_program(synth.retinaface_state_dict(), fused=False)
    kinds = [op.type for op in P.ops]
    assert kinds.count(nat.TR_OP_STEM) == 1 and kinds.count(nat.TR_OP_DWCONV) == 13
    # 56 convs = 1 stem + 13 depthwise + 33 dense + 9 heads fused into 3
    assert kinds.count(nat.TR_OP_CONV) == 56 - 1 - 13 - 9 + 3
    assert len(roles['heads']) == 3
    # fused program: every depthwise lives inside the 1x1 conv that consumes it
    Pf, rf = weights.retinaface_program(synth.retinaface_state_dict(), fused=True)
    kinds = [op.type for op in Pf.ops]
    assert kinds.count(nat.TR_OP_SEPCONV) == 13 and kinds.count(nat.TR_OP_DWCONV) == 0
    assert kinds.count(nat.TR_OP_CONV) == 56 - 1 - 13 - 13 - 9 + 3
    seps = [op for op in Pf.ops if op.type == nat.TR_OP_SEPCONV]
    assert [(o.cin_real, o.cout_real, o.stride) for o in seps] == [
        (8, 16, 1), (16, 32, 2), (32, 32, 1), (32, 64, 2), (64, 64, 1), (64, 128, 2)] + \
        [(128, 128, 1)] * 5 + [(128, 256, 2), (256, 256, 1)]
    assert len(rf['heads']) == 3
    P, _ = weights.openpose_program(synth.openpose_state_dict())
    kinds = [op.type for op in P.ops]
    # 92 convs; in every stage the layers of the two branches are ONE op each (the first reads a
    # shared input with the filter banks stacked, the middle ones are groups = 2 convs, the last
    # is one conv over the pair whose rows are the concat channels): 12 trunk + 5 + 5 * 7 launches
    assert kinds.count(nat.TR_OP_CONV) + kinds.count(nat.TR_OP_STEM) == 92 - 5 - 6 * 5 - 5
    assert sum(op.groups == 2 for op in P.ops) == 3 + 5 * 5
    assert kinds.count(nat.TR_OP_MAXPOOL) == 3
    small = (1, 1, 1, 1)
    P, _ = weights.arcface_program(synth.arcface_state_dict(units=small), units=small)
    kinds = [op.type for op in P.ops]
    # stem + per unit (2 convs + shortcut for first units) + FC
    assert kinds.count(nat.TR_OP_CONV) == 4 * 3 + 1 and kinds.count(nat.TR_OP_VIEW) == 1
    # concat-channel remap for the 7x7 stage convs
    m = weights.OPENPOSE_CAT_MAP
    assert len(m) == 185 and m[37] == 37 and m[38] == 40 and m[56] == 58 and m[57] == 64 and m[-1] == 191


def test_umeyama_recovers_similarity():
    from oracle.align import umeyama_similarity, LANDMARK_TEMPLATE
    th, s, t = 0.3, 1.7, np.array([5.0, -3.0])
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    src = LANDMARK_TEMPLATE.astype(np.float64)
    dst = (s * (R @ src.T)).T + t
    T = umeyama_similarity(src, dst)
    np.testing.assert_allclose(T[:2, :2], s * R, atol=1e-9)
    np.testing.assert_allclose(T[:2, 2], t, atol=1e-9)


def test_closed_form_similarity_matches_umeyama():
    """The product's closed-form 2-D similarity (what ``tr_face_similarity`` also computes on
    the device) against the oracle's SVD Umeyama + matrix inverse on 10 000 random faces:
    rotated / scaled / translated / jittered templates, mirrored ones included."""
    from oracle.align import alignment_coefficients, LANDMARK_TEMPLATE
    from terran_b200.face.recognition.arcface.wrapper import similarity_coefficients
    rng = np.random.default_rng(3)
    n = 10000
    th = rng.uniform(-np.pi, np.pi, n)
    sc = rng.uniform(0.2, 8.0, n)
    t = rng.uniform(-500, 1500, (n, 1, 2))
    R = np.stack([np.stack([np.cos(th), -np.sin(th)], -1), np.stack([np.sin(th), np.cos(th)], -1)], -2)
    base = LANDMARK_TEMPLATE.astype(np.float64)[None] * np.where(rng.random(n) < 0.1, -1.0, 1.0)[:, None, None] ** np.array([1, 0])
    lm = (sc[:, None, None] * np.einsum('nij,nkj->nki', R, base) + t + rng.normal(0, 1.5, (n, 5, 2))).astype(np.float32)
    got = similarity_coefficients(lm)
    want = np.stack([alignment_coefficients(l) for l in lm])
    rel = np.abs(got - want) / np.maximum(1e-3, np.abs(want))
    assert np.isfinite(got).all() and rel.max() < 1e-9, rel.max()


# -------------------------------------------------------------- multi-process

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import torch.distributed as dist
from terran_b200 import parallel, synth
rank, world, _ = parallel.init_from_env(backend='gloo')
assert world == 2
# 1. weight broadcast: rank 1 starts from nothing
sd = synth.retinaface_state_dict() if rank == 0 else None
sd = parallel.broadcast_state_dict(sd, src=0)
ref = synth.retinaface_state_dict()
assert list(sd) == list(ref)
for k in ref:
    assert sd[k].dtype == ref[k].dtype and torch.equal(sd[k], ref[k]), k
# 2. frame sharding: a pure per-frame function, sharded == unsharded
frames = np.arange(7 * 3, dtype=np.float32).reshape(7, 3)
def call(batch):        # stand-in for a model call: variable number of rows per frame
    return [np.full((int(f[0]) % 4, 16), f.sum(), np.float32) for f in batch]
lo, hi, res = parallel.sharded_call(call, frames)
assert (lo, hi) == parallel.shard_range(7, rank, 2)
counts = [len(r) for r in res]
rows = np.concatenate(res) if res else np.zeros((0, 16), np.float32)
out = parallel.gather_rows(counts, rows, dst=0)
if rank == 0:
    want = call(frames)
    assert len(out) == 7
    for a, b in zip(out, want):
        assert a.shape == b.shape and np.array_equal(a, b)
    print('GATHER_OK')
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_broadcast_and_gather(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    r = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29731', str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'GATHER_OK' in r.stdout


def test_shard_range_partitions():
    from terran_b200.parallel import shard_range
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the unmodified reference from baseline/_ref — or, where it
    is not installed, the oracle port — timed on host cores) prints ONE JSON line with the
    arm's keys: same metric / unit / config as the GPU arm, a cpu_baseline describing the run
    and an e2e block without device traffic."""
    import json
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='4')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--warmup', '0', '--ref-frames', '2'], capture_output=True,
                       text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('frames/sec end-to-end (detect+pose)')
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1
    from baseline import refarm
    want_kind = 'reference' if refarm.reference_path() else 'port'
    assert d['cpu_baseline']['kind'] == want_kind and d['cpu_baseline']['cores'] >= 1
    assert d['config']['frames_per_gpu'] == 2 and d['config']['people_per_frame'] > 0
    assert d['cpu_baseline']['value'] == d['value'] and 'sample' in d['cpu_baseline']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}


def test_resample_tables_of_the_library_match_the_oracle():
    """tr_resample_table (host-only entry of the C ABI: the fixed-point filter tables the
    letterbox kernels read) == oracle/letterbox.py::resample_table, which is pinned against PIL."""
    from oracle.letterbox import resample_table
    from terran_b200 import _native as nat
    L = nat.lib()
    rng = np.random.default_rng(2)
    cases = [(int(a), int(b)) for a, b in zip(rng.integers(1, 3000, 120), rng.integers(1, 160, 120))]
    cases += [(112, 112), (1, 1), (1, 112), (1920, 112), (1080, 63), (5, 112)]
    for in_size, out_size in cases:
        ksize = L.tr_resample_table(in_size, out_size, None, None)
        bounds = np.zeros((out_size, 2), np.int32)
        coeffs = np.full((out_size, ksize), -7, np.int32)
        assert L.tr_resample_table(in_size, out_size, bounds.ctypes.data, coeffs.ctypes.data) == ksize
        want_b, want_k = resample_table(in_size, out_size)
        assert want_k.shape[1] == ksize
        np.testing.assert_array_equal(bounds, want_b)
        np.testing.assert_array_equal(coeffs, want_k)
    assert L.tr_resample_table(0, 5, None, None) == -1
    sizes = np.array([[1, 500]], np.int32)            # resized height 0: Pillow's own error
    assert L.tr_face_letterbox_workspace_bytes(sizes.ctypes.data, 1, 112) == 0
    assert b'must be > 0' in L.tr_last_error()
    sizes = np.array([[300, 181], [112, 112]], np.int32)
    assert L.tr_face_letterbox_workspace_bytes(sizes.ctypes.data, 2, 112) > 300 * 67 * 3


def test_result_waits_block_when_ranks_crowd_the_cpus(monkeypatch):
    """``defaults.blocking_waits``: spinning waits (the CUDA default) unless the ranks of the box
    reach a quarter of the CPUs this process may run on; the module switch and TRB_BLOCKING_SYNC
    override the rule (profiles/r02_e2e_ranks.txt: same window time, 4.5x less CPU when blocking)."""
    import os
    from terran_b200 import defaults
    cpus = len(os.sched_getaffinity(0))
    monkeypatch.delenv('TRB_BLOCKING_SYNC', raising=False)
    monkeypatch.setattr(defaults, 'blocking_events', None)
    monkeypatch.setenv('LOCAL_WORLD_SIZE', '1')
    assert defaults.blocking_waits() == (4 >= cpus)
    monkeypatch.setenv('LOCAL_WORLD_SIZE', str(cpus))
    assert defaults.blocking_waits() is True
    monkeypatch.setenv('TRB_BLOCKING_SYNC', '0')
    assert defaults.blocking_waits() is False
    monkeypatch.setenv('TRB_BLOCKING_SYNC', '1')
    monkeypatch.setenv('LOCAL_WORLD_SIZE', '1')
    assert defaults.blocking_waits() is True
    monkeypatch.setattr(defaults, 'blocking_events', False)
    assert defaults.blocking_waits() is False
