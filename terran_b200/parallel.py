"""Frame-sharded data parallelism: one process per GPU, no collective on the
per-frame step.

The reference is single-process / single-device; every
frame is independent on its hot path, so the only communication this adds is
  * one broadcast of the checkpoint tensors from rank 0 at start-up, and
  * an optional variable-length gather of per-frame results to rank 0
    (all-gather of per-frame counts, then one padded all-gather of the rows).
Works over NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world_size):
    """Contiguous frame range [lo, hi) of ``rank``: sizes differ by at most one
    and the concatenation over ranks is range(n_frames)."""
    lo = rank * n_frames // world_size
    hi = (rank + 1) * n_frames // world_size
    return lo, hi


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment.  Returns
    (rank, world_size, local_rank); a no-op single-process world when
    WORLD_SIZE is absent or 1."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gpu_numa_cpus(device_index):
    """(numa node, cpu ids) of the NUMA node the GPU's PCIe root hangs off, read from sysfs
    (``/sys/bus/pci/devices/<bus id>/numa_node`` and ``/sys/devices/system/node/nodeN/cpulist``);
    (None, None) when the platform does not say."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node'
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None, None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            spec = f.read().strip()
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None, None
    cpus = []
    for part in spec.split(','):
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return node, cpus


def bind_to_gpu_numa(device_index, local_world=1):
    """Pin this process to the CPUs of its GPU's NUMA node — its own slice of them when several
    of the ``local_world`` ranks (GPU i = local rank i) share the node — BEFORE it allocates
    pinned staging buffers, so that the host->device copies of every rank read local memory
    instead of all ranks streaming from node 0.  Returns a dict describing the binding
    (reported by bench.py); a no-op when the topology is unknown or the mask cannot change."""
    node, cpus = gpu_numa_cpus(device_index)
    info = {'numa_node': node, 'cpus': None}
    if not cpus:
        return info
    peers = [i for i in range(max(local_world, device_index + 1)) if gpu_numa_cpus(i)[0] == node]
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        share = max(1, len(allowed) // max(1, len(peers)))
        slot = peers.index(device_index) if device_index in peers else 0
        mine = allowed[slot * share:(slot + 1) * share] or allowed
        os.sched_setaffinity(0, mine)
        info['cpus'] = f'{mine[0]}-{mine[-1]} ({len(mine)})'
    except (OSError, AttributeError):
        pass
    return info


def _comm_device():
    if dist.is_initialized() and dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def broadcast_state_dict(state_dict, src=0):
    """Broadcast a checkpoint from ``src``: every float tensor travels in ONE
    flat fp32 buffer (a single collective), integer buffers in a second one.
    Ranks other than ``src`` pass a state_dict with the right keys/shapes (e.g.
    a freshly initialised one) or ``None`` to receive keys and shapes too."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return state_dict
    meta = [None]
    if dist.get_rank() == src:
        meta[0] = [(k, tuple(v.shape), str(v.dtype)) for k, v in state_dict.items()]
    dist.broadcast_object_list(meta, src=src)
    dev = _comm_device()
    out = {}
    for floats in (True, False):
        items = [m for m in meta[0] if m[2].startswith('torch.float') == floats]
        if not items:
            continue
        dtype = torch.float32 if floats else torch.int64
        total = sum(int(np.prod(s)) for _, s, _ in items)
        flat = torch.empty(total, dtype=dtype, device=dev)
        if dist.get_rank() == src:
            flat.copy_(torch.cat([state_dict[k].reshape(-1).to(dtype) for k, _, _ in items]))
        dist.broadcast(flat, src=src)
        flat = flat.cpu()
        off = 0
        for k, shape, dt in items:
            n = int(np.prod(shape))
            out[k] = flat[off:off + n].reshape(shape).to(getattr(torch, dt.split('.')[1])).clone()
            off += n
    return {k: out[k] for k, _, _ in meta[0]}


def gather_rows(counts, rows, dst=0):
    """Variable-length gather.  ``counts``: (n_local,) int — rows per local
    frame; ``rows``: (sum(counts), D) float32.  Returns on ``dst`` the list over
    ALL frames (rank order = frame order) of (count_i, D) arrays; None elsewhere."""
    counts = np.asarray(counts, np.int64)
    rows = np.asarray(rows, np.float32).reshape(int(counts.sum()), -1)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.split(rows, np.cumsum(counts)[:-1]) if len(counts) else []
    world, rank, dev = dist.get_world_size(), dist.get_rank(), _comm_device()
    D = torch.tensor([len(counts), rows.shape[0], rows.shape[1]], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(D) for _ in range(world)]
    dist.all_gather(sizes, D)
    sizes = torch.stack(sizes).cpu().numpy()
    max_frames, max_rows, width = int(sizes[:, 0].max()), int(sizes[:, 1].max()), int(sizes[:, 2].max())
    c_pad = torch.zeros(max(max_frames, 1), dtype=torch.int64, device=dev)
    c_pad[:len(counts)] = torch.from_numpy(counts).to(dev)
    r_pad = torch.zeros((max(max_rows, 1), max(width, 1)), dtype=torch.float32, device=dev)
    if rows.size:
        r_pad[:rows.shape[0], :rows.shape[1]] = torch.from_numpy(rows).to(dev)
    all_c = [torch.empty_like(c_pad) for _ in range(world)]
    all_r = [torch.empty_like(r_pad) for _ in range(world)]
    dist.all_gather(all_c, c_pad)
    dist.all_gather(all_r, r_pad)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        nf, nr = int(sizes[r, 0]), int(sizes[r, 1])
        c = all_c[r][:nf].cpu().numpy()
        rr = all_r[r][:nr, :width].cpu().numpy()
        out += np.split(rr, np.cumsum(c)[:-1]) if nf else []
    return out


def gather_faces(faces_per_frame, dst=0):
    """Gather ``Detection`` results of this rank's frame shard to ``dst`` in frame order.
    faces_per_frame: list (local frames) of lists of face dicts (``bbox`` int32 (4,),
    ``landmarks`` int32 (5,2), ``score`` float32).  Returns the list over ALL frames on ``dst``
    (None elsewhere).  Pixel coordinates travel as float32, exact below 2**24."""
    counts = [len(f) for f in faces_per_frame]
    rows = np.zeros((sum(counts), 15), np.float32)
    i = 0
    for faces in faces_per_frame:
        for f in faces:
            rows[i, :4] = f['bbox']
            rows[i, 4:14] = np.asarray(f['landmarks']).reshape(-1)
            rows[i, 14] = f['score']
            i += 1
    frames = gather_rows(counts, rows, dst=dst)
    if frames is None:
        return None
    return [[{'bbox': r[:4].astype(np.int32), 'landmarks': r[4:14].astype(np.int32).reshape(5, 2),
              'score': np.float32(r[14])} for r in fr] for fr in frames]


def track_sharded(tracker, faces_per_frame, dst=0):
    """Face tracking needs frame ORDER (caveat): the detections of every rank's
    shard are gathered to ``dst`` and fed to ``tracker`` (a ``terran_b200.tracking.Sort``) one
    frame after the other there.  Returns the tracked faces per frame on ``dst``, None elsewhere."""
    frames = gather_faces(faces_per_frame, dst=dst)
    if frames is None:
        return None
    return [tracker.update(faces) for faces in frames]


def sharded_call(fn, frames, rank=None, world_size=None):
    """Apply ``fn`` (a model ``call``) to this rank's contiguous shard of
    ``frames``; returns (lo, hi, results)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(frames), rank, world_size)
    return lo, hi, (fn(frames[lo:hi]) if hi > lo else [])
