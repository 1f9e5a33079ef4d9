"""DMD on the POD basis with the reference's signatures (pyLOM/DMD/wrapper.py:50-146).

Everything that touches the m-row arrays runs on the device through the C ABI: centering, the TSQR-SVD of the first
n-1 snapshots, the projection U^T Y2 (transposed-tall DMMA GEMM + one all-reduce, the reference's `matmulp`), the
mode product Phi = Y2 (V S^-1 w / mu) and the reconstruction (tall DMMA GEMMs; a complex m x N array is read and
written through its interleaved real view, so no re/im split passes).  The r x r / r x n complex algebra between
them -- eig of Atilde, Vandermonde, Cholesky, two triangular inverses -- is control-plane sized and done with
numpy on the host, which is also where the reference does it (`eigen` is CPU-only there, vmmath/maths.py:160-182).
"""
import numpy as np
import torch

from .. import _dev
from ..utils.cr import cr, cr_start, cr_stop
from ..utils import parall
from ..vmmath.averaging import temporal_mean, subtract_mean
from ..vmmath.maths import matmul, matmul_tn
from ..vmmath.svd import _tsqr_svd_dev, _engine as _svd_engine
from ..vmmath.truncation import compute_truncation_residual


def _host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _out(x, kind, device):
    """Small results follow the array kind of the input (numpy in -> numpy out, device tensor in -> device tensor)."""
    if kind.endswith("|f32"):            # float32 caller: small results in single precision too
        kind = kind[:-4]
        x = np.asarray(x).astype(np.complex64 if np.iscomplexobj(x) else np.float32)
    if kind == "numpy":
        return x
    t = torch.from_numpy(np.ascontiguousarray(x))
    return t if kind == "torch_cpu" else t.to(device)


def _vandermonde(mu, n):
    """Vand[:, k] = mu**k (pyLOM/vmmath/maths.py:202-222)."""
    Vand = np.zeros((mu.shape[0], n), dtype=np.complex128)
    for icol in range(n):
        Vand[:, icol] = mu ** icol
    return Vand


def _interleaved(M):
    """Complex (k x N) -> real (k x 2N) with columns re0, im0, re1, im1, ...: a real GEMM with this B operand writes
    its (m x 2N) result in exactly the memory layout of a complex128 (m x N) array."""
    B = np.empty((M.shape[0], 2 * M.shape[1]))
    B[:, 0::2] = M.real
    B[:, 1::2] = M.imag
    return B


def _order_modes(muReal, muImag, Phi, bJov):
    """Sort by decreasing |b| and put the positive-imaginary member of every conjugate pair first
    (pyLOM/DMD/wrapper.py:18-47, statement for statement; Phi is the device array, mu / b live on the host)."""
    order = np.flip(np.abs(bJov).argsort()).copy()
    muReal, muImag, bJov = muReal[order], muImag[order], bJov[order]
    Phi = Phi[:, torch.from_numpy(order).to(Phi.device)]
    Pri = torch.view_as_real(Phi)            # (m, N, 2) view: [..., 1] is Phi.imag
    p = False
    for ii in range(muImag.shape[0]):
        if p:
            p = False
            continue
        iimag = muImag[ii]
        if iimag < 0:
            muImag[ii] = muImag[ii + 1]
            muImag[ii + 1] = -muImag[ii]
            bJov.imag[ii] = bJov.imag[ii + 1]
            bJov.imag[ii + 1] = -bJov.imag[ii]
            Pri[:, ii, 1] = Pri[:, ii + 1, 1]
            Pri[:, ii + 1, 1] = -Pri[:, ii + 1, 1]
            p = True
            continue
        if iimag > 0:
            p = True
            continue
    return muReal, muImag, Phi, bJov


@cr('DMD.run')
def run(X, r, remove_mean=True):
    """DMD of the (row-distributed) snapshot matrix X(m_i, n).

    Returns muReal (N), muImag (N), Phi (m_i, N) complex modes, bJov (N) complex amplitudes (Jovanovic et al. 2014),
    ordered by decreasing |b| (pyLOM/DMD/wrapper.py:50-117).  X is not modified.
    """
    Xd, kind = _dev.to_device(X, "X")
    if remove_mean:
        cr_start('DMD.temporal_mean', 0)
        Y = subtract_mean(Xd, temporal_mean(Xd))
        cr_stop('DMD.temporal_mean', 0)
    else:
        Y = Xd
    muReal, muImag, Phi, bJov = _run_dev(Y, r)
    dev = Y.device
    return _out(muReal, kind, dev), _out(muImag, kind, dev), _dev.from_device(Phi, kind), _out(bJov, kind, dev)


def _run_dev(Y, r, engine=None):
    """DMD of the (centred) device matrix Y; `engine` supplies the device operations (tests inject a CPU stand-in to
    run the multi-rank composition over gloo)."""
    eng = engine or _svd_engine
    m, n = Y.shape
    if n < 2:
        raise ValueError("DMD.run needs at least two snapshots")
    cr_start('DMD.SVD', 0)
    U, S, VT, _ = _tsqr_svd_dev(Y[:, :-1].contiguous(), engine=engine)
    cr_stop('DMD.SVD', 0)
    N = int(r) if r >= 1 else compute_truncation_residual(S, r)
    U, S, VT = U[:, :N], S[:N], VT[:N, :]

    # Atilde = U^T Y2 V S^-1 with Y2 = Y[:, 1:].  U^T Y is formed over all n columns (16-byte aligned rows) and the
    # first column dropped: one more column of flops instead of the unaligned-load path.
    cr_start('DMD.linear_mapping', 0)
    aux1 = parall.mpi_reduce(eng.matmul_tn(U, Y), op='sum', all=True)[:, 1:]
    S_h, VT_h, aux1_h = _host(S), _host(VT), _host(aux1)
    Atilde = aux1_h @ (VT_h / S_h[:, None]).T
    cr_stop('DMD.linear_mapping', 0)

    cr_start('DMD.modes', 0)
    mu, w = np.linalg.eig(Atilde)
    muReal, muImag = np.real(mu).copy(), np.imag(mu).copy()
    M = ((VT_h.T * (1.0 / S_h)) @ w) / mu                       # (n-1, N): V S^-1 w / mu
    Bm = np.vstack((np.zeros((1, 2 * N)), _interleaved(M)))     # zero row for the dropped first snapshot
    Phi = torch.view_as_complex(eng.matmul(Y, torch.from_numpy(Bm).to(Y.device)).contiguous().view(m, N, 2))
    cr_stop('DMD.modes', 0)

    cr_start('DMD.amplitudes', 0)
    Vand = _vandermonde(mu, n - 1)
    P = (w.conj().T @ w) * np.conj(Vand @ Vand.conj().T)
    Pl = np.linalg.cholesky(P)
    G = S_h[:, None] * VT_h
    q = np.conj(np.diag((Vand @ G.conj().T) @ w))
    bJov = np.linalg.inv(Pl.conj().T) @ (np.linalg.inv(Pl) @ q)
    cr_stop('DMD.amplitudes', 0)

    cr_start('DMD.order', 0)
    muReal, muImag, Phi, bJov = _order_modes(muReal, muImag, Phi, bJov)
    cr_stop('DMD.order', 0)
    return muReal, muImag, Phi, bJov


@cr('DMD.frequency_damping')
def frequency_damping(real, imag, dt):
    """Damping ratio log|mu|/dt and frequency arg(mu)/dt of every mode (pyLOM/DMD/wrapper.py:119-130)."""
    if isinstance(real, torch.Tensor):
        return torch.log(torch.sqrt(real * real + imag * imag)) / dt, torch.atan2(imag, real) / dt
    real, imag = np.asarray(real), np.asarray(imag)
    return np.log(np.sqrt(real * real + imag * imag)) / dt, np.arctan2(imag, real) / dt


@cr('DMD.mode_computation')
def mode_computation(X, V, S, W):
    """X V^T S^-1 |W| (pyLOM/DMD/wrapper.py:132-138): the small factors are multiplied first, the tall product is one
    DMMA GEMM."""
    Xd, kind = _dev.to_device(X, "X")
    small = (_host(V).T * (1.0 / _host(S))) @ np.abs(_host(W))
    return _dev.from_device(matmul(Xd, torch.from_numpy(np.ascontiguousarray(small)).to(Xd.device)), kind)


@cr('DMD.reconstruction_jovanovic')
def reconstruction_jovanovic(Phi, real, imag, t, bJov):
    """Re(Phi diag(b) Vand(t)) (pyLOM/DMD/wrapper.py:140-146), Vand[:, it] = mu**t[it] (vmmath/maths.py:224-246)."""
    kind = "torch"
    if isinstance(Phi, np.ndarray):
        Phi, kind = torch.from_numpy(np.ascontiguousarray(Phi)), "numpy"
    if not Phi.is_cuda:
        _dev.require_cuda()
        kind = "numpy" if kind == "numpy" else "torch_cpu"
        Phi = Phi.cuda()
    if Phi.dtype != torch.complex128:
        raise NotImplementedError("only complex128 modes are implemented on the B200 path")
    Phi = Phi.contiguous()
    m, N = Phi.shape
    mu = _host(real) + 1j * _host(imag)
    th = _host(t)
    Vand = np.zeros((N, th.shape[0]), dtype=np.complex128)
    for it, tt in enumerate(th):
        Vand[:, it] = mu ** tt
    C = _host(bJov)[:, None] * Vand                              # (N, nt) = diag(b) Vand
    Br = np.empty((2 * N, th.shape[0]))
    Br[0::2] = C.real                                           # Re(Phi C) = Phi_re C_re - Phi_im C_im
    Br[1::2] = -C.imag
    Xr = matmul(torch.view_as_real(Phi).view(m, 2 * N), torch.from_numpy(Br).to(Phi.device))
    return _dev.from_device(Xr, kind)
