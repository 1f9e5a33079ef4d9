"""POD.run / truncate / reconstruct with the reference's signatures (pyLOM/POD/wrapper.py:16-103,
compiled twin pyLOM/POD/wrapper.pyx:95-377)."""
import torch

from .. import _lib, _dev
from ..utils.cr import cr, cr_start, cr_stop
from ..vmmath.svd import _tsqr_svd_dev, randomized_svd
from ..vmmath.truncation import compute_truncation_residual
from ..vmmath.averaging import temporal_mean, subtract_mean, temporal_variance, norm_variance


@cr('POD.run')
def run(X, remove_mean=True, divide_variance=False, randomized=False, r=1, q=3, seed=-1):
    """POD of the (row-distributed) snapshot matrix X(m_i, n).

    Returns U (m_i, n) spatial modes, S (n) singular values, V (n, n) = V^T temporal coefficients.
    X is not modified.  The centering (temporal_mean + subtract_mean, POD/wrapper.py:33-41) is fused
    into the copy that feeds the factorisation.  With randomized=True the r leading modes come from
    randomized_svd (q power iterations, sketch seeded with `seed`): U (m_i, r), S (r), V (r, n).
    """
    Xd, kind = _dev.to_device(X, "X")
    center = bool(remove_mean)
    if remove_mean and (divide_variance or randomized):
        cr_start('POD.temporal_mean', 0)
        X_mean = temporal_mean(Xd)
        if divide_variance:                  # POD/wrapper.py:36-38 (only effective together with remove_mean)
            Xd = norm_variance(Xd, X_mean, temporal_variance(Xd, X_mean))
        else:                                # the randomized path reads Y several times: materialise it once
            Xd = subtract_mean(Xd, X_mean)
        cr_stop('POD.temporal_mean', 0)
        center = False
    cr_start('POD.SVD', 0)
    if randomized:
        U, S, V = randomized_svd(Xd, r, q, seed=seed)
    else:
        U, S, V, _ = _tsqr_svd_dev(Xd, center=center)
    cr_stop('POD.SVD', 0)
    return _dev.from_device(U, kind), _dev.from_device(S, kind), _dev.from_device(V, kind)


@cr('POD.truncate')
def truncate(U, S, V, r=1e-8):
    """Keep N modes: r >= 1 -> N = int(r); 0 < r < 1 -> residual target; r < 0 -> cumulative energy
    (POD/wrapper.py:55-82).  Returns views U[:, :N], S[:N], V[:N, :] like the reference's .py path."""
    N = int(r) if r >= 1 else compute_truncation_residual(S, r)
    return U[:, :N], S[:N], V[:N, :]


@cr('POD.reconstruct')
def reconstruct(U, S, V):
    """X(m_i, n) = U(m_i, N) diag(S) V(N, n); the temporal mean is NOT re-added (POD/wrapper.py:86-103)."""
    kind = "torch"
    keep = lambda t: isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64   # fp64 device views: zero copy
    if not keep(U):
        U, kind = _dev.to_device(U, "U")
    Sd, _ = _dev.to_device(S, "S")
    Vd = V if keep(V) else _dev.to_device(V, "V")[0]
    if U.stride(1) != 1 and U.shape[1] > 1:
        U = U.contiguous()
    if Vd.stride(1) != 1 and Vd.shape[1] > 1:
        Vd = Vd.contiguous()
    m, N = U.shape
    N2, n = Vd.shape
    if N != N2 or Sd.numel() != N:
        raise ValueError("reconstruct: inconsistent shapes")
    ldu = U.stride(0) if m > 1 else N
    ldv = Vd.stride(0) if N > 1 else n
    X = torch.empty((m, n), dtype=torch.float64, device=U.device)
    L = _lib.lib()
    _, wp, wb = _dev.workspace(L.pl_matmul_workspace_bytes(n, N), "matmul", U.device)
    _lib.check(L.pl_reconstruct_f64(X.data_ptr(), U.data_ptr(), ldu, Sd.data_ptr(), Vd.data_ptr(), ldv, m, N, n, wp, wb,
                                    _dev.stream()), "reconstruct")
    return _dev.from_device(X, kind)
