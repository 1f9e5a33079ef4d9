"""temporal_mean / subtract_mean (pyLOM/vmmath/averaging.py:17-44, src/averaging.c:29-46,109-124)."""
import torch

from .. import _lib, _dev
from ..utils.cr import cr


@cr('math.temporal_mean')
def temporal_mean(X):
    """Temporal mean of X(m,n): one value per row (m = spatial dofs, n = snapshots)."""
    Xd, kind = _dev.to_device(X, "X")
    m, n = Xd.shape
    out = torch.empty(m, dtype=torch.float64, device=Xd.device)
    _lib.check(_lib.lib().pl_temporal_mean_f64(out.data_ptr(), Xd.data_ptr(), m, n, _dev.stream()), "temporal_mean")
    return _dev.from_device(out, kind)


@cr('math.subtract_mean')
def subtract_mean(X, X_mean):
    """out(m,n) = X(m,n) - X_mean(m)."""
    Xd, kind = _dev.to_device(X, "X")
    Md, _ = _dev.to_device(X_mean, "X_mean")
    m, n = Xd.shape
    if Md.numel() != m:
        raise ValueError("X_mean must have one entry per row of X")
    out = torch.empty_like(Xd)
    _lib.check(_lib.lib().pl_subtract_mean_f64(out.data_ptr(), Xd.data_ptr(), Md.data_ptr(), m, n, _dev.stream()),
               "subtract_mean")
    return _dev.from_device(out, kind)


@cr('math.temporal_variance')
def temporal_variance(X, X_mean):
    """Population variance of every row, shape (m, 1) like `p.var(X, axis=1, keepdims=True)`
    (pyLOM/vmmath/averaging.py:46-59, src/averaging.c:70-90)."""
    Xd, kind = _dev.to_device(X, "X")
    Md, _ = _dev.to_device(X_mean, "X_mean")
    m, n = Xd.shape
    out = torch.empty(m, dtype=torch.float64, device=Xd.device)
    _lib.check(_lib.lib().pl_temporal_variance_f64(out.data_ptr(), Xd.data_ptr(), Md.reshape(-1).data_ptr(), m, n, _dev.stream()),
               "temporal_variance")
    return _dev.from_device(out.reshape(m, 1), kind)


@cr('math.temporal_variance')
def norm_variance(X, X_mean, X_var):
    """(X - X_mean) / X_var  (pyLOM/vmmath/averaging.py:61-74, src/averaging.c:109-158)."""
    Xd, kind = _dev.to_device(X, "X")
    Md, _ = _dev.to_device(X_mean, "X_mean")
    Vd, _ = _dev.to_device(X_var, "X_var")
    m, n = Xd.shape
    out = torch.empty_like(Xd)
    _lib.check(_lib.lib().pl_norm_variance_f64(out.data_ptr(), Xd.data_ptr(), Md.reshape(-1).data_ptr(),
                                              Vd.reshape(-1).data_ptr(), m, n, _dev.stream()), "norm_variance")
    return _dev.from_device(out, kind)


@cr('math.center')
def center(X):
    """Fused temporal_mean + subtract_mean: returns (Y, X_mean).  Not in the reference API; it is what
    POD.run does back to back (POD/wrapper.py:33-41)."""
    Xd, kind = _dev.to_device(X, "X")
    m, n = Xd.shape
    Y = torch.empty_like(Xd)
    mean = torch.empty(m, dtype=torch.float64, device=Xd.device)
    _lib.check(_lib.lib().pl_center_f64(Y.data_ptr(), mean.data_ptr(), Xd.data_ptr(), m, n, _dev.stream()), "center")
    return _dev.from_device(Y, kind), _dev.from_device(mean, kind)
