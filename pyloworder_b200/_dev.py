"""Device-buffer plumbing: torch owns memory and streams, nothing else."""
import numpy as np
import torch

from . import _lib

_ws_cache = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pyloworder_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def to_device(x, what="array"):
    """Return (fp64 contiguous CUDA tensor, kind) where kind says how to hand results back."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float64:
            raise NotImplementedError(f"{what}: only float64 is implemented on the B200 path (got {x.dtype})")
        if not x.is_cuda:
            return x.cuda(non_blocking=True).contiguous(), "torch_cpu"
        return x.contiguous(), "torch"
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64:
            raise NotImplementedError(f"{what}: only float64 is implemented on the B200 path (got {x.dtype})")
        return torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=True), "numpy"
    if hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
        if t.dtype != torch.float64:
            raise NotImplementedError(f"{what}: only float64 is implemented on the B200 path")
        return t.contiguous(), "cai"
    raise TypeError(f"{what}: unsupported array type {type(x)}")


def from_device(t, kind):
    if kind == "numpy":
        return t.cpu().numpy()
    if kind == "torch_cpu":
        return t.cpu()
    return t


def workspace(nbytes, tag, device):
    """Cached uint8 scratch tensor (torch's caching allocator owns the memory)."""
    key = (tag, device)
    w = _ws_cache.get(key)
    if w is None or w.numel() < nbytes + 256:
        _ws_cache.pop(key, None)
        del w                          # release the old buffer BEFORE allocating the larger one (each is ~ the matrix size)
        w = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = w
    off = (-w.data_ptr()) % 256
    return w, w.data_ptr() + off, w.numel() - off


def cached_workspace_bytes(tag, device):
    w = _ws_cache.get((tag, device))
    return 0 if w is None else w.numel()


def free_workspaces():
    """Release every cached workspace (they are otherwise kept, grow-only, for the next call of the same shape)."""
    _ws_cache.clear()


def ptr(t):
    return 0 if t is None else t.data_ptr()
