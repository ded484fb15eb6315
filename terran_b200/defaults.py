"""Default device (reference ``terran/defaults.py:3-5``).  The B200 classes have
no CPU path: constructing a model on a CPU device raises."""
import torch

default_device = (
    torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
)


def cuda_index(device):
    """CUDA ordinal of ``device`` (torch.device / str / int); raises on CPU."""
    device = torch.device(device) if not isinstance(device, torch.device) else device
    if device.type != 'cuda':
        raise RuntimeError(
            'terran_b200 runs on a B200 GPU only: got device '
            f'{device} (there is no CPU fallback)')
    return device.index if device.index is not None else torch.cuda.current_device()


#: How host threads wait for results: False = the CUDA default (the waiting thread spins on the
#: event, lowest latency), True = blocking waits (the thread sleeps until the driver's interrupt).
#: ``None`` (default) decides per process in ``completion_event``: blocking when the ranks of
#: this box reach a quarter of the CPUs they may run on — spinning waits of many ranks then
#: take the cores the feeder threads and the result unpacking need.
blocking_events = None


def blocking_waits():
    """Whether result waits of this process should block (see ``blocking_events``)."""
    import os
    if blocking_events is not None:
        return bool(blocking_events)
    env = os.environ.get('TRB_BLOCKING_SYNC')
    if env is not None:
        return env not in ('0', '')
    ranks = int(os.environ.get('LOCAL_WORLD_SIZE', '1'))
    try:
        cpus = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cpus = os.cpu_count() or 1
    return ranks * 4 >= cpus


def completion_event():
    """Event a host thread will ``synchronize()`` on for the results of a batch."""
    return torch.cuda.Event(blocking=blocking_waits())
