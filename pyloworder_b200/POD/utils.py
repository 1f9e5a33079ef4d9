"""POD helpers next to the hot path (pyLOM/POD/utils.py:19-42)."""
import numpy as np
import torch


def extract_modes(U, ivar, npoints, modes=[], reshape=True):
    """Separate the spatial modes of variable `ivar` (1-based) when several variables were concatenated per point
    (rows ivar-1, ivar-1+nvars, ... of U).  Pure indexing: works on device tensors and numpy arrays alike.

    Returns (len(modes)*npoints,) if reshape else (npoints, len(modes)), like the reference.
    """
    nvars = U.shape[0] // npoints
    if len(modes) == 0:
        modes = list(range(1, U.shape[1] + 1))
    cols = [int(m) - 1 for m in modes]
    rows = slice(ivar - 1, nvars * npoints, nvars)
    if isinstance(U, torch.Tensor):
        out = U[rows][:, torch.as_tensor(cols, device=U.device)].contiguous()
        return out.reshape(len(cols) * npoints) if reshape else out
    out = np.ascontiguousarray(U[rows][:, cols])
    return out.reshape((len(cols) * npoints,), order='C') if reshape else out
