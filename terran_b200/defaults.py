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
