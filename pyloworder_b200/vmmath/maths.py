"""matmul / vecmat / matmulp (pyLOM/vmmath/maths.py:76-129, src/vector_matrix.c:206-242,344-356,401-414)."""
import torch

from .. import _lib, _dev
from ..utils.cr import cr
from ..utils.parall import mpi_reduce


def _strided2d(t):
    """Accept row-major tensors whose rows are contiguous (views from POD.truncate)."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D array")
    if t.stride(1) != 1 and t.shape[1] > 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous(); ld = t.shape[1]
    return t, ld


def _to_dev_keep_view(x, what):
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64:
        return x, "torch"
    return _dev.to_device(x, what)


@cr('math.matmul')
def matmul(A, B):
    """C(M,N) = A(M,Q) x B(Q,N) on the FP64 tensor cores."""
    Ad, kind = _to_dev_keep_view(A, "A")
    Bd, _ = _to_dev_keep_view(B, "B")
    Ad, lda = _strided2d(Ad)
    Bd, ldb = _strided2d(Bd)
    m, k = Ad.shape
    k2, n = Bd.shape
    if k != k2:
        raise ValueError(f"matmul: inner dimensions differ ({k} vs {k2})")
    C = torch.empty((m, n), dtype=torch.float64, device=Ad.device)
    L = _lib.lib()
    _, wp, wb = _dev.workspace(L.pl_matmul_workspace_bytes(n, k), "matmul", Ad.device)
    _lib.check(L.pl_matmul_f64(C.data_ptr(), n, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, m, n, k, wp, wb, _dev.stream()),
               "matmul")
    return _dev.from_device(C, kind)


def _tall_transposed(A, what):
    """A is (M, Q) with Q the long, row-distributed dimension (callers pass `Ai.T` / `Qi.T`).  Return the row-major
    (Q, M) operand the kernel reads, without materialising a transposed copy when A is already a `.T` view."""
    if isinstance(A, torch.Tensor) and A.is_cuda:
        if A.dtype != torch.float64:
            raise NotImplementedError(f"{what}: only float64 is implemented on the B200 path (got {A.dtype})")
        if A.dim() != 2:
            raise ValueError("expected a 2-D array")
        return A.T, "torch"            # _strided2d() below copies only if the rows of A.T are not contiguous
    if hasattr(A, "T") and not hasattr(A, "__cuda_array_interface__"):
        return _dev.to_device(A.T, what)          # numpy / CPU torch: uploads A.T (no host copy for a .T view)
    t, kind = _dev.to_device(A, what)
    return t.T, kind


@cr('math.matmul_tn')
def matmul_tn(X, Y):
    """C(a, b) = X^T Y for two row-major tall device operands X (m, a), Y (m, b): the rank-local part of matmulp
    (transposed-tall DMMA GEMM, reduction over the rows)."""
    X, ldx = _strided2d(X)
    Y, ldy = _strided2d(Y)
    m, a = X.shape
    m2, b = Y.shape
    if m != m2:
        raise ValueError(f"matmulp: inner dimensions differ ({m} vs {m2})")
    C = torch.empty((a, b), dtype=torch.float64, device=X.device)
    L = _lib.lib()
    _, wp, wb = _dev.workspace(L.pl_matmul_tn_workspace_bytes(a, b), "matmul_tn", X.device)
    _lib.check(L.pl_matmul_tn_f64(C.data_ptr(), b, X.data_ptr(), ldx, a, Y.data_ptr(), ldy, b, m, wp, wb, _dev.stream()),
               "matmul_tn")
    return C


@cr('math.matmulp')
def matmulp(A, B):
    """C(M,N) = A(M,Q) x B(Q,N) with Q the row-distributed dimension; the result is summed over the ranks and is the
    same on all of them (pyLOM/vmmath/maths.py:93-110).  pyLOM calls it as matmulp(Ai.T, Qi) / matmulp(Qi.T, Ai):
    both operands are tall row-major arrays and the product is a reduction over their rows."""
    X, kind = _tall_transposed(A, "A")
    Bd, _ = _to_dev_keep_view(B, "B")
    C = matmul_tn(X, Bd)
    return _dev.from_device(mpi_reduce(C, op='sum', all=True), kind)


@cr('math.vecmat')
def vecmat(v, A):
    """C[i,:] = v[i] * A[i,:]."""
    vd, _ = _dev.to_device(v, "v")
    Ad, kind = _dev.to_device(A, "A")
    m, n = Ad.shape
    C = torch.empty_like(Ad)
    _lib.check(_lib.lib().pl_vecmat_f64(C.data_ptr(), vd.data_ptr(), Ad.data_ptr(), m, n, _dev.stream()), "vecmat")
    return _dev.from_device(C, kind)


def vector_sum(v, start=0):
    """Sum of v[start:] (pyLOM/vmmath/maths.py:32-44) -- O(n) scalar work on the n singular values."""
    return float(torch.as_tensor(v)[start:].sum())


def vector_norm(v, start=0):
    """2-norm of v[start:] (pyLOM/vmmath/maths.py:47-59)."""
    return float(torch.linalg.vector_norm(torch.as_tensor(v)[start:]))
