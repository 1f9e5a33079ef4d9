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


_F32 = "|f32"      # suffix of `kind`: the caller's arrays are float32 (results are handed back as float32)


def _widen(t32):
    """float32 CUDA tensor -> float64 CUDA tensor through the library's streaming kernel (pl_widen_f32_f64)."""
    t32 = t32.contiguous()
    out = torch.empty(t32.shape, dtype=torch.float64, device=t32.device)
    if t32.numel():
        with torch.cuda.device(t32.device):
            _lib.check(_lib.lib().pl_widen_f32_f64(out.data_ptr(), t32.data_ptr(), t32.numel(), stream(t32.device)), "widen_f32_f64")
    return out


def _narrow(t64):
    t64 = t64.contiguous()
    out = torch.empty(t64.shape, dtype=torch.float32, device=t64.device)
    if t64.numel():
        with torch.cuda.device(t64.device):
            _lib.check(_lib.lib().pl_narrow_f64_f32(out.data_ptr(), t64.data_ptr(), t64.numel(), stream(t64.device)), "narrow_f64_f32")
    return out


def to_device(x, what="array"):
    """Return (fp64 contiguous CUDA tensor, kind) where kind says how to hand results back.

    float32 inputs (the reference's `real` fused type also covers float: stsqr_svd, pyLOM/vmmath/src/svd.c:529-563) are
    widened to fp64 on the device, the fp64 path runs, and `from_device` narrows the results again: fp32 callers get
    fp32 arrays back.  Complex inputs (SPOD) are not implemented."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.float32:
            kind = ("torch" if x.is_cuda else "torch_cpu") + _F32
            return _widen(x if x.is_cuda else x.cuda(non_blocking=True)), kind
        if x.dtype != torch.float64:
            raise NotImplementedError(f"{what}: only float64 / float32 are implemented on the B200 path (got {x.dtype})")
        if not x.is_cuda:
            return x.cuda(non_blocking=True).contiguous(), "torch_cpu"
        return x.contiguous(), "torch"
    if isinstance(x, np.ndarray):
        if x.dtype == np.float32:
            return _widen(torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=True)), "numpy" + _F32
        if x.dtype != np.float64:
            raise NotImplementedError(f"{what}: only float64 / float32 are implemented on the B200 path (got {x.dtype})")
        return torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=True), "numpy"
    if hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
        if t.dtype == torch.float32:
            return _widen(t), "cai" + _F32
        if t.dtype != torch.float64:
            raise NotImplementedError(f"{what}: only float64 / float32 are implemented on the B200 path")
        return t.contiguous(), "cai"
    raise TypeError(f"{what}: unsupported array type {type(x)}")


def from_device(t, kind):
    if kind.endswith(_F32):
        kind = kind[:-len(_F32)]
        if isinstance(t, torch.Tensor) and t.dtype == torch.float64:
            t = _narrow(t)
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
