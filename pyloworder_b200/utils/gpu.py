"""Device selection / transfers (pyLOM/utils/gpu.py:18-23,49-66) on torch instead of cupy."""
import numpy as np
import torch

from .parall import rank


def gpu_device(id=None, gpu_per_node=4):
    """Select device `id % gpu_per_node` (default id = rank), pyLOM/utils/gpu.py:18-23."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device")
    n = min(gpu_per_node, torch.cuda.device_count())
    dev = (rank() if id is None else id) % n
    torch.cuda.set_device(dev)
    return dev


def gpu_to_cpu(X):
    return X.cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)


def cpu_to_gpu(X):
    return torch.as_tensor(X).cuda()
