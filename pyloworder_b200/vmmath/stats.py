"""RMSE (pyLOM/vmmath/stats.py:17-34, src/stats.c:44-72) and the reconstruction energy
(pyLOM/vmmath/truncation.py:40-60): fused sum((A-B)^2), sum(A^2) kernel + one 2-scalar all-reduce."""
import numpy as np
import torch

from .. import _lib, _dev
from ..utils.cr import cr
from ..utils.parall import mpi_reduce, is_distributed


def _diff_sums(A, B, what):
    """Device tensor [sum((A-B)^2), sum(A^2)] over the local rows (one streaming pass over both arrays)."""
    Ad, _ = _dev.to_device(A, "A")
    Bd, _ = _dev.to_device(B, "B")
    if Ad.shape != Bd.shape:
        raise ValueError(f"{what}: shapes differ")
    L = _lib.lib()
    out = torch.empty(2, dtype=torch.float64, device=Ad.device)
    _, wp, _ = _dev.workspace(L.pl_rmse_workspace_bytes(), "rmse", Ad.device)
    _lib.check(L.pl_rmse_sums_f64(out.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), Ad.numel(), wp, _dev.stream()), what)
    return out, Ad.numel()


@cr('math.RMSE')
def RMSE(A, B, relative=True):
    out, count = _diff_sums(A, B, "RMSE")
    if not relative:
        out[1] = float(count)
    if is_distributed():
        out = mpi_reduce(out, op='sum', all=True)
    s = out.cpu().numpy()
    return float(np.sqrt(s[0] / s[1]))


@cr('math.energy')
def energy(original, rec):
    """Reconstruction energy 1 - sum((original - rec)^2) / sum(original^2), sums over all ranks
    (pyLOM/vmmath/truncation.py:40-60)."""
    out, _ = _diff_sums(original, rec, "energy")
    if is_distributed():
        out = mpi_reduce(out, op='sum', all=True)
    s = out.cpu().numpy()
    return float(1.0 - s[0] / s[1])
