"""Plugin registry: task/alias -> model class, class -> local weight file.

Mirrors the lookup half of the reference's ``terran/checkpoint.py`` (``CHECKPOINTS``
:29-103, ``get_terran_home`` :106-126, ``read_checkpoint_db`` :145-169,
``get_checkpoint`` :172-210, ``get_class_for_checkpoint`` :213-245,
``get_checkpoint_path`` :277-328) with the same ids, aliases, home directory
(``~/.terran`` or ``$TERRAN_HOME``) and ``ValueError('Checkpoint not found.')``
behaviour, so reference ``.pth`` files are picked up unchanged.  The
downloader and the click CLI are out of scope (no per-frame work, no network).

``register_with_reference()`` appends these classes to an installed reference's
``terran.checkpoint.CHECKPOINTS`` under the alias ``'b200'`` so that the
unmodified reference wrappers resolve ``Detection(checkpoint='b200')`` to the
B200 classes.
"""
import importlib
import os
from pathlib import Path

__all__ = ['get_terran_home', 'get_class_for_checkpoint', 'get_checkpoint_path']

DEFAULT_TERRAN_HOME = Path('~/.terran')
CHECKPOINT_PATH = 'checkpoints'

CHECKPOINTS = [
    {
        'id': 'b5d77fff',
        'name': 'RetinaFace',
        'description': 'RetinaFace with mnet backbone (B200-native kernels).',
        'task': 'face-detection',
        'class': 'terran_b200.face.detection.retinaface.RetinaFace',
        'alias': 'gpu-realtime',
        'default': True,
    },
    {
        'id': 'd206e4b0',
        'name': 'ArcFace',
        'description': 'ArcFace with Resnet 100 backbone (B200-native kernels).',
        'task': 'face-recognition',
        'class': 'terran_b200.face.recognition.arcface.ArcFace',
        'alias': 'gpu-realtime',
        'default': True,
    },
    {
        'id': '11a769ad',
        'name': 'OpenPose',
        'description': 'OpenPose with VGG backend, 2017 version (B200-native kernels).',
        'task': 'pose-estimation',
        'class': 'terran_b200.pose.openpose.OpenPose',
        'alias': 'gpu-realtime',
        'default': True,
    },
]


def get_terran_home(create_if_missing=True):
    path = Path(os.environ.get('TERRAN_HOME', DEFAULT_TERRAN_HOME)).expanduser()
    if create_if_missing and not path.exists():
        path.mkdir(exist_ok=True)
    return path


def get_checkpoints_directory():
    path = get_terran_home() / CHECKPOINT_PATH
    path.mkdir(exist_ok=True)
    return path


def read_checkpoint_db():
    local = set(p.stem for p in get_checkpoints_directory().glob('*.pth'))
    return {'checkpoints': [
        {
            'status': 'DOWNLOADED' if c['id'] in local else 'NOT_DOWNLOADED',
            'local_path': (get_checkpoints_directory() / f"{c['id']}.pth"
                           if c['id'] in local else None),
            **c,
        }
        for c in CHECKPOINTS
    ]}


def get_checkpoint(db, id_or_alias):
    if isinstance(id_or_alias, tuple):
        task_name, alias = id_or_alias
        selected = [
            c for c in db['checkpoints']
            if c['task'] == task_name and (
                c['alias'] == alias if alias is not None else c['default'])
        ]
    else:
        selected = [c for c in db['checkpoints'] if c['id'] == id_or_alias]
    return selected[0] if selected else None


def get_class_for_checkpoint(task_name, alias):
    checkpoint = get_checkpoint(read_checkpoint_db(), (task_name, alias))
    if not checkpoint:
        raise ValueError('Checkpoint not found.')
    module_path, class_name = checkpoint['class'].rsplit('.', maxsplit=1)
    return getattr(importlib.import_module(module_path), class_name)


def get_checkpoint_path(model_class_path):
    """Local ``.pth`` of the model class; there is no downloader here, so a
    checkpoint that is not on disk is an error."""
    db = read_checkpoint_db()
    selected = [c for c in db['checkpoints'] if c['class'] == model_class_path]
    if not selected:
        raise ValueError('Checkpoint not found.')
    checkpoint = selected[0]
    if checkpoint['status'] == 'NOT_DOWNLOADED':
        raise ValueError(
            f"Checkpoint '{checkpoint['id']}' is not present under "
            f"{get_checkpoints_directory()} (no network access from here: place the "
            f"reference's .pth there, or pass `state_dict=` to the model class)."
        )
    return checkpoint['local_path']


def register_with_reference(alias='b200'):
    """Make an installed reference resolve ``checkpoint=alias`` to these classes."""
    import terran.checkpoint as ref   # noqa: the reference package, if installed
    for c in CHECKPOINTS:
        entry = dict(c, alias=alias, default=False)
        if not any(e['class'] == entry['class'] and e['alias'] == alias for e in ref.CHECKPOINTS):
            ref.CHECKPOINTS.append(entry)
