"""RMSE (pyLOM/vmmath/stats.py:17-34, src/stats.c:44-72): fused sum((A-B)^2), sum(A^2) kernel + one
2-scalar all-reduce."""
import numpy as np
import torch

from .. import _lib, _dev
from ..utils.cr import cr
from ..utils.parall import mpi_reduce, is_distributed


@cr('math.RMSE')
def RMSE(A, B, relative=True):
    Ad, _ = _dev.to_device(A, "A")
    Bd, _ = _dev.to_device(B, "B")
    if Ad.shape != Bd.shape:
        raise ValueError("RMSE: shapes differ")
    L = _lib.lib()
    out = torch.empty(2, dtype=torch.float64, device=Ad.device)
    _, wp, _ = _dev.workspace(L.pl_rmse_workspace_bytes(), "rmse", Ad.device)
    _lib.check(L.pl_rmse_sums_f64(out.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), Ad.numel(), wp, _dev.stream()), "RMSE")
    if not relative:
        out[1] = float(Ad.numel())
    if is_distributed():
        out = mpi_reduce(out, op='sum', all=True)
    s = out.cpu().numpy()
    return float(np.sqrt(s[0] / s[1]))
