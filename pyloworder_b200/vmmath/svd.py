"""qr / svd / tsqr / tsqr_svd -- pyLOM/vmmath/svd.py:28-118,227-252 (src/svd.c:83-139,280-321,565-712).

Single rank: one C call (`pl_tsqr_svd_f64`).  P ranks (one process per GPU, torch.distributed):

    R_i        = local Householder QR (CUDA)                      <- dqr at svd.c:594
    Rstack     = ONE all-gather of the n x n R_i (NCCL/NVLink)    <- replaces the butterfly svd.c:602-669
    R, Q2      = Householder QR of the (P n) x n stack, redundantly on every GPU (bit-identical)
    Ur, S, VT  = Jacobi SVD of R (CUDA)                           <- dsvd at svd.c:706
    U_i        = Q1_i (Q2_i Ur)                                   <- dmatmul at svd.c:673 and :708 fused

TSQR is valid for any reduction tree (Demmel et al. 2012, the paper svd.py:58 cites), so the flat
all-gather tree gives the same R up to row signs.  S and VT are identical on all ranks.
"""
import os
import time

import numpy as _np
import torch

from .. import _lib, _dev
from ..utils.cr import cr, cr_start, cr_stop
from ..utils import parall


def next_power_of_2(n):
    """pyLOM/vmmath/svd.py:17-25."""
    p = 1
    if n and not (n & (n - 1)):
        return n
    while p < n:
        p <<= 1
    return p


def _inplace_ok(m, n, device):
    """Use the in-place variant (output buffer = factorisation buffer, ~2.3 x A instead of ~3.4 x A, one extra
    device-to-device pass over U)?  Needs n % 32 == 0; chosen when the out-of-place buffers would not fit in
    the free device memory (PL_INPLACE=1 / PL_NO_INPLACE=1 force the choice)."""
    if n % 32 != 0 or os.environ.get("PL_NO_INPLACE"):
        return False
    if os.environ.get("PL_INPLACE"):
        return True
    need = _lib.lib().pl_qr_workspace_bytes(m, n) + m * n * 8
    free, _ = torch.cuda.mem_get_info(device)
    free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    free += _dev.cached_workspace_bytes("local", device)      # an already cached workspace is reused, not allocated again
    return need > 0.92 * free


class CudaEngine:
    """The device operations the multi-rank composition needs; tests inject a CPU stand-in."""

    def __init__(self):
        self._ubuf = {}      # tag -> (m + n + 32) x n buffer that holds the reflectors / Q / U of the in-place path
        self._side = {}      # device -> side stream of the multi-rank path

    def factor(self, A, tag, center=False):
        m, n = A.shape
        L = _lib.lib()
        R = torch.empty((n, n), dtype=torch.float64, device=A.device)
        mean = torch.empty(m, dtype=torch.float64, device=A.device) if center else None
        if _inplace_ok(m, n, A.device):
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes_inplace(m, n), tag, A.device)
            ubuf = torch.empty((L.pl_qr_inplace_rows(m, n), n), dtype=torch.float64, device=A.device)
            _lib.check(L.pl_qr_factor_inplace_f64(R.data_ptr(), _dev.ptr(mean), ubuf.data_ptr(), A.data_ptr(), m, n, int(center),
                                                  wp, wb, _dev.stream()), "qr_factor")
            self._ubuf[tag] = ubuf
        else:
            self._ubuf.pop(tag, None)
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes(m, n), tag, A.device)
            _lib.check(L.pl_qr_factor_f64(R.data_ptr(), _dev.ptr(mean), A.data_ptr(), m, n, int(center), wp, wb, _dev.stream()),
                       "qr_factor")
        return R, mean

    def form_q(self, shape, tag, device):
        """Turn the reflectors of the matrix last factored under `tag` into the explicit Q1 (in its workspace);
        a following apply_q(..., formed=True) only multiplies.  Needs nothing but the local factorisation, so the
        multi-rank path runs it while the R factors are exchanged and the small SVD is computed on another stream."""
        m, n = shape
        L = _lib.lib()
        ubuf = self._ubuf.get(tag)
        if ubuf is not None:
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes_inplace(m, n), tag, device)
            _lib.check(L.pl_qr_apply_q_inplace_f64(ubuf.data_ptr(), 0, 0, m, n, 2, wp, wb, _dev.stream()), "qr_form_q")
        else:
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes(m, n), tag, device)
            _lib.check(L.pl_qr_apply_q_f64(0, n, 0, 0, n, m, n, 2, wp, wb, _dev.stream()), "qr_form_q")

    def apply_q(self, shape, W, tag, device, formed=False):
        """U = Q1 W for the matrix last factored under `tag` (W None -> explicit Q1)."""
        m, n = shape
        L = _lib.lib()
        flags = 1 if formed else 0
        ubuf = self._ubuf.pop(tag, None)
        if ubuf is not None:
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes_inplace(m, n), tag, device)
            if W is not None and (W.shape[0] != n or W.shape[1] != n):
                raise ValueError("apply_q: W must be n x n")
            ldw = 0 if W is None else W.stride(0)
            _lib.check(L.pl_qr_apply_q_inplace_f64(ubuf.data_ptr(), _dev.ptr(W), ldw, m, n, flags, wp, wb, _dev.stream()), "qr_apply_q")
            return ubuf[:m]
        _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes(m, n), tag, device)
        nw = n if W is None else W.shape[1]
        U = torch.empty((m, nw), dtype=torch.float64, device=device)
        ldw = 0 if W is None else W.stride(0)
        _lib.check(L.pl_qr_apply_q_f64(U.data_ptr(), nw, _dev.ptr(W), ldw, nw, m, n, flags, wp, wb, _dev.stream()), "qr_apply_q")
        return U

    def side_stream(self, device):
        """High-priority stream for the exchange + small factorisations of the multi-rank path (None: run in line)."""
        if os.environ.get("PL_NO_SVD_OVERLAP"):
            return None
        s = self._side.get(device)
        if s is None:
            s = self._side[device] = torch.cuda.Stream(device=device, priority=-1)
        return s

    def svd(self, R):
        n = R.shape[0]
        L = _lib.lib()
        _, wp, wb = _dev.workspace(L.pl_svd_workspace_bytes(n), "svd", R.device)
        U = torch.empty((n, n), dtype=torch.float64, device=R.device)
        S = torch.empty(n, dtype=torch.float64, device=R.device)
        VT = torch.empty((n, n), dtype=torch.float64, device=R.device)
        _lib.check(L.pl_svd_f64(U.data_ptr(), S.data_ptr(), VT.data_ptr(), R.data_ptr(), n, wp, wb, _dev.stream()), "svd")
        return U, S, VT

    def tsqr_svd_single(self, A, center=False):
        m, n = A.shape
        L = _lib.lib()
        S = torch.empty(n, dtype=torch.float64, device=A.device)
        VT = torch.empty((n, n), dtype=torch.float64, device=A.device)
        mean = torch.empty(m, dtype=torch.float64, device=A.device) if center else None
        if _inplace_ok(m, n, A.device):
            _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes_inplace(m, n), "local", A.device)
            ubuf = torch.empty((L.pl_qr_inplace_rows(m, n), n), dtype=torch.float64, device=A.device)
            _lib.check(L.pl_pod_run_inplace_f64(ubuf.data_ptr(), S.data_ptr(), VT.data_ptr(), _dev.ptr(mean), A.data_ptr(), m, n,
                                                int(center), wp, wb, _dev.stream()), "tsqr_svd")
            return ubuf[:m], S, VT, mean
        _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes(m, n), "local", A.device)
        U = torch.empty((m, n), dtype=torch.float64, device=A.device)
        _lib.check(L.pl_pod_run_f64(U.data_ptr(), S.data_ptr(), VT.data_ptr(), _dev.ptr(mean), A.data_ptr(), m, n,
                                    int(center), wp, wb, _dev.stream()), "tsqr_svd")
        return U, S, VT, mean

    def tsqr_svd_dist(self, A, center=False):
        """P ranks, ONE collective C call (pl_tsqr_svd_dist_f64): local QR, ncclAllGather of R inside the library,
        redundant stack QR + Jacobi, back-multiply -- the counterpart of the reference's collective dtsqr_svd."""
        m, n = A.shape
        L = _lib.lib()
        comm = parall.c_comm()
        S = torch.empty(n, dtype=torch.float64, device=A.device)
        VT = torch.empty((n, n), dtype=torch.float64, device=A.device)
        mean = torch.empty(m, dtype=torch.float64, device=A.device) if center else None
        inplace = _inplace_ok(m, n, A.device)
        flags = 1 if inplace else 0
        _, wp, wb = _dev.workspace(L.pl_tsqr_svd_dist_workspace_bytes(comm, m, n, flags), "local", A.device)
        rows = L.pl_qr_inplace_rows(m, n) if inplace else m
        U = torch.empty((rows, n), dtype=torch.float64, device=A.device)
        _lib.check(L.pl_tsqr_svd_dist_f64(comm, U.data_ptr(), S.data_ptr(), VT.data_ptr(), _dev.ptr(mean), A.data_ptr(), m, n,
                                          int(center), flags, wp, wb, _dev.stream()), "tsqr_svd_dist")
        return U[:m], S, VT, mean

    def allgather_rows(self, R):
        return parall.mpi_allgather_rows(R)

    def matmul(self, A, B):
        from .maths import matmul
        return matmul(A, B)

    def matmul_tn(self, X, Y):
        from .maths import matmul_tn
        return matmul_tn(X, Y)

    def svd_any(self, A):
        return _svd_dev(A)


_engine = CudaEngine()


def _check_shape(A, collective=False):
    """Shape precondition of the TSQR path.  With `collective` the verdict is all-reduced first, so that every rank
    raises together instead of one rank raising while the others wait in the exchange."""
    if A.dim() != 2:
        raise ValueError("expected a 2-D array (m, n)")
    m, n = A.shape
    bad = m < n
    if collective and parall.is_distributed():
        bad = bool(parall.mpi_reduce(1.0 if bad else 0.0, op="max") > 0)
    if bad:
        raise ValueError(f"every rank needs at least n rows (this rank: m_i={m}, n={n}); "
                         "the reference has the same precondition (pyLOM/vmmath/svd.py:69)")


def _tsqr_svd_dev(Ad, center=False, engine=None):
    """Core composition on device tensors.  Returns (U_i, S, VT, mean_i)."""
    eng = engine or _engine
    _check_shape(Ad, collective=True)
    P, rank = parall.size(), parall.rank()
    if P == 1:
        return eng.tsqr_svd_single(Ad, center)
    m, n = Ad.shape
    if engine is None and parall.use_c_comm():
        return eng.tsqr_svd_dist(Ad, center)
    R_i, mean = eng.factor(Ad, "local", center)

    def small_part():
        cr_start('math.tsqr.allgather')
        Rstack = eng.allgather_rows(R_i)
        cr_stop('math.tsqr.allgather')
        R, _ = eng.factor(Rstack, "stack", False)
        Ur, S, VT = eng.svd(R)
        return eng.apply_q((P * n, n), Ur, "stack", Ad.device), S, VT        # Q2 Ur, (P n) x n, tiny

    side = eng.side_stream(Ad.device) if hasattr(eng, "side_stream") and m >= 1_000_000 else None
    if side is None:
        Wfull, S, VT = small_part()
        U = eng.apply_q((m, n), Wfull[rank * n:(rank + 1) * n], "local", Ad.device)
        return U, S, VT, mean
    # The explicit Q1 needs only the local reflectors: it is formed on the main stream while the all-gather, the QR of
    # the stack and the latency-bound Jacobi SVD run on a high-priority side stream (same overlap as on one rank).
    main = torch.cuda.current_stream(Ad.device)
    side.wait_stream(main)
    eng.form_q((m, n), "local", Ad.device)
    with torch.cuda.stream(side):
        Wfull, S, VT = small_part()
        R_i.record_stream(side)                  # allocated on the main stream, read on the side stream
        for t in (Wfull, S, VT):                 # allocated on the side stream, read on the main stream
            t.record_stream(main)
    main.wait_stream(side)
    U = eng.apply_q((m, n), Wfull[rank * n:(rank + 1) * n], "local", Ad.device, formed=True)
    return U, S, VT, mean


def _complex_select(S2, V2, n):
    """Host part of the complex path (2n numbers / a 2n x 2n matrix).  The real embedding doubles every singular value;
    the right vectors [c; d] of a cluster of 2k equal values span, as complex vectors c + i d, a k-dimensional complex
    space (v and i v are both in it).  Per cluster keep the k vectors picked by pivoted complex Gram-Schmidt (largest
    residual first), then orthonormalise the kept set: the correction M is the identity up to rounding unless complex
    singular values coincide (k > 1: repeated values, several zeros).  Returns the kept indices J (n, ascending), M
    (n x n complex, None when |M - I| <= 1e-12) and the orthonormal complex right vectors as the columns of X."""
    Vc = V2[:, :n] + 1j * V2[:, n:]                 # row j: v_j^T as a complex row
    tol = 1e-10 * max(float(S2[0]), 1e-300)
    J = []
    j0 = 0
    while j0 < 2 * n:
        j1 = j0 + 1
        while j1 < 2 * n and (abs(S2[j1] - S2[j0]) <= tol or (j1 - j0) % 2 == 1):   # clusters have an even number of members
            j1 += 1
        k = (j1 - j0) // 2
        cand = Vc[j0:j1].copy()                      # residuals, updated in place
        for _ in range(k):
            nrm = _np.einsum("ij,ij->i", cand.conj(), cand).real
            p = int(_np.argmax(nrm))
            J.append(j0 + p)
            q = cand[p] / _np.sqrt(nrm[p])
            cand = cand - _np.outer(cand @ q.conj(), q)
            cand[p] = 0.0
        j0 = j1
    J = _np.array(sorted(J))
    if len(J) != n:
        raise RuntimeError("complex tsqr_svd: could not separate the doubled singular pairs of the real embedding")
    X = Vc[J].T                                      # columns v_k
    Qx, Rx = _np.linalg.qr(X)
    ph = _np.diag(Rx) / _np.abs(_np.diag(Rx))        # keep the phases of the kept vectors
    M = _np.linalg.inv(Rx) * ph[None, :]
    if _np.abs(M - _np.eye(n)).max() <= 1e-12:
        return J, None, X
    return J, M, Qx * ph[None, :]


def _tsqr_svd_complex(Ac):
    """complex128 tsqr_svd (ztsqr_svd, pyLOM/vmmath/src/svd.c:955-1010; the call of SPOD, pyLOM/SPOD/wrapper.py:83): the fp64 path on
    the real embedding Ahat = [[Ar, -Ai], [Ai, Ar]] of the local shard (a row permutation of the global embedding, which
    does not change S, V and permutes U accordingly).  Ahat = Qhat Rhat (TSQR, explicit Qhat), Rhat = Ur diag(S2) V2^T
    (Jacobi, 2n x 2n), every complex singular triplet appears twice; keep one member per pair:
    U = Qhat Ur[:, J] read as top + i bottom, S = S2[J], V^H rows = conj(c + i d) of the kept rows [c; d] of V2^T.
    Costs twice the flops and memory of a native complex factorisation."""
    L = _lib.lib()
    Ac = Ac.contiguous()
    m, n = Ac.shape
    Ahat = torch.empty((2 * m, 2 * n), dtype=torch.float64, device=Ac.device)
    _lib.check(L.pl_complex_embed_f64(Ahat.data_ptr(), Ac.data_ptr(), m, n, _dev.stream()), "complex_embed")
    Qh, Rh = _tsqr_dev(Ahat)
    del Ahat
    Urh, S2, VT2 = _engine.svd(Rh.contiguous())
    J, M, X = _complex_select(S2.cpu().numpy(), VT2.cpu().numpy(), n)
    Jd = torch.from_numpy(J).to(Ac.device)
    W = Urh.index_select(1, Jd).contiguous()                     # (2n, n) column gather of the small factor
    U = torch.empty((m, n), dtype=torch.complex128, device=Ac.device)
    if M is None:
        P = _engine.matmul(Qh, W)
        _lib.check(L.pl_complex_pack_f64(U.data_ptr(), P.data_ptr(), 0, m, n, _dev.stream()), "complex_pack")
    else:                                                        # coinciding complex singular values: U <- U M
        Md = torch.from_numpy(_np.ascontiguousarray(M)).to(Ac.device)
        Uh = _engine.matmul(Qh, W)                               # [Ur; Ui]
        P = _engine.matmul(Uh, Md.real.contiguous())
        Q = _engine.matmul(Uh, Md.imag.contiguous())
        _lib.check(L.pl_complex_pack_f64(U.data_ptr(), P.data_ptr(), Q.data_ptr(), m, n, _dev.stream()), "complex_pack")
    S = S2.index_select(0, Jd).contiguous()
    VH = torch.from_numpy(_np.ascontiguousarray(X.conj().T)).to(Ac.device)     # rows v_k^H: A = U diag(S) VH
    return U, S, VH


@cr('math.tsqr_svd')
def tsqr_svd(Ai):
    """SVD of the row-distributed matrix via TSQR.  Ai(m_i,n) -> Ui(m_i,n), S(n), V(n,n) = V^T (V^H for complex input)."""
    if (isinstance(Ai, torch.Tensor) and Ai.dtype == torch.complex128) or (isinstance(Ai, _np.ndarray) and Ai.dtype == _np.complex128):
        _dev.require_cuda()
        is_np = isinstance(Ai, _np.ndarray)
        Ac = torch.from_numpy(_np.ascontiguousarray(Ai)) if is_np else Ai
        on_host = not Ac.is_cuda
        if Ac.dim() != 2:
            raise ValueError("expected a 2-D array (m, n)")
        U, S, VH = _tsqr_svd_complex(Ac.cuda() if on_host else Ac)
        if is_np:
            return U.cpu().numpy(), S.cpu().numpy(), VH.cpu().numpy()
        return (U.cpu(), S.cpu(), VH.cpu()) if on_host else (U, S, VH)
    Ad, kind = _dev.to_device(Ai, "Ai")
    U, S, VT, _ = _tsqr_svd_dev(Ad)
    return _dev.from_device(U, kind), _dev.from_device(S, kind), _dev.from_device(VT, kind)


def _tsqr_dev(Ad, engine=None):
    """Q_i (m_i, n), R (n, n) on device tensors: local Householder QR, one all-gather of the R_i, redundant QR of the stack."""
    eng = engine or _engine
    _check_shape(Ad)
    P, rank = parall.size(), parall.rank()
    m, n = Ad.shape
    R_i, _ = eng.factor(Ad, "local")
    if P == 1:
        return eng.apply_q((m, n), None, "local", Ad.device), R_i
    Rstack = eng.allgather_rows(R_i)
    R, _ = eng.factor(Rstack, "stack")
    Q2 = eng.apply_q((P * n, n), None, "stack", Ad.device)
    return eng.apply_q((m, n), Q2[rank * n:(rank + 1) * n], "local", Ad.device), R


@cr('math.tsqr')
def tsqr(Ai):
    """Parallel QR: Qi(m_i,n), R(n,n) identical on all ranks (pyLOM/vmmath/svd.py:49-118)."""
    Ad, kind = _dev.to_device(Ai, "Ai")
    Q, R = _tsqr_dev(Ad)
    return _dev.from_device(Q, kind), _dev.from_device(R, kind)


@cr('math.qr')
def qr(A):
    """Thin Householder QR of a local matrix: Q(m,n), R(n,n) (pyLOM/vmmath/svd.py:28-36)."""
    Ad, kind = _dev.to_device(A, "A")
    _check_shape(Ad)
    m, n = Ad.shape
    R, _ = _engine.factor(Ad, "local")
    Q = _engine.apply_q((m, n), None, "local", Ad.device)
    return _dev.from_device(Q, kind), _dev.from_device(R, kind)


def _svd_dev(Ad):
    """Thin SVD of a local device matrix of any shape: square -> Jacobi kernel; tall -> the single-rank TSQR-SVD;
    wide (r x n, the B of randomized_svd) -> TSQR-SVD of the transpose with the factors swapped back."""
    if Ad.dim() != 2:
        raise ValueError("expected a 2-D array")
    m, n = Ad.shape
    if m == n:
        return _engine.svd(Ad.contiguous())
    if m > n:
        U, S, VT, _ = _engine.tsqr_svd_single(Ad.contiguous())
        return U, S, VT
    Ut, S, VTt, _ = _engine.tsqr_svd_single(Ad.T.contiguous())        # A^T = Ut S VTt  =>  A = VTt^T S Ut^T
    return VTt.T.contiguous(), S, Ut.T.contiguous()


@cr('math.svd')
def svd(A, method='gesdd'):
    """Thin SVD of a local (not distributed) matrix: U (m,k), S (k, descending), V^T (k,n), k = min(m,n)
    (pyLOM/vmmath/svd.py:38-47, dsvd src/svd.c:83-139).  `method` is accepted for signature compatibility."""
    Ad, kind = _dev.to_device(A, "A")
    U, S, VT = _svd_dev(Ad)
    return _dev.from_device(U, kind), _dev.from_device(S, kind), _dev.from_device(VT, kind)


def _sketch_matrix(n, r, rng, device):
    """omega = rand(n, r) from numpy's MT19937: the same numbers the reference draws with
    `np.random.seed(seed); np.random.rand(n, r)` (pyLOM/vmmath/svd.py:131-133), generated on the host (n*r values)
    from a private RandomState, i.e. without touching numpy's global generator."""
    return torch.from_numpy(rng.rand(int(n), int(r))).to(device)


def _power_sketch(Ad, omega, q, eng):
    """Y = A (A^T A)^q omega with a TSQR re-orthonormalisation between the products (the loop at svd.py:134-140).
    `matmulp(Ai.T, Qi)` of the reference = the transposed-tall product X^T Y over the local rows + one all-reduce."""
    Yi = eng.matmul(Ad, omega)
    for _ in range(int(q)):
        Qi, _R = _tsqr_dev(Yi, eng)
        Q2i = parall.mpi_reduce(eng.matmul_tn(Ad, Qi), op='sum', all=True)      # (n, r) = A^T Q
        Yi = eng.matmul(Ad, Q2i)
    return Yi


def _check_sketch_shape(m, n, r):
    r = int(r)
    if not 1 <= r <= n:
        raise ValueError(f"randomized_qr: need 1 <= r <= n (got r={r}, n={n})")
    if m < r:
        raise ValueError(f"every rank needs at least r rows (got m_i={m} < r={r})")
    return r


def _randomized_qr_dev(Ad, r, q, seed, engine=None, rng=None):
    """randomized_qr on device tensors: returns Q_i (m_i, r), B (r, n) = Q^T A and the sketch Y_i."""
    eng = engine or _engine
    m, n = Ad.shape
    r = _check_sketch_shape(m, n, r)
    if rng is None:
        rng = _np.random.RandomState(int(time.time()) if seed is None or seed < 0 else int(seed))
    omega = _sketch_matrix(n, r, rng, Ad.device)
    Yi = _power_sketch(Ad, omega, q, eng)
    Qi, _R = _tsqr_dev(Yi, eng)
    B = parall.mpi_reduce(eng.matmul_tn(Qi, Ad), op='sum', all=True)            # (r, n) = Q^T A
    return Qi, B, Yi


def _randomized_svd_dev(Ad, r, q, seed, engine=None):
    eng = engine or _engine
    Qi, B, _Y = _randomized_qr_dev(Ad, r, q, seed, eng)
    Ur, S, V = eng.svd_any(B)
    return eng.matmul(Qi, Ur), S, V


@cr('math.randomized_qr')
def randomized_qr(Ai, r, q, seed=-1):
    """Randomized range finder (pyLOM/vmmath/svd.py:120-144; drandomized_qr src/svd.c:1267-1319):
    Ai (m_i, n) -> Qi (m_i, r) orthonormal over all ranks, B (r, n) = Q^T A identical on all ranks."""
    Ad, kind = _dev.to_device(Ai, "Ai")
    Qi, B, _Y = _randomized_qr_dev(Ad, r, q, seed)
    return _dev.from_device(Qi, kind), _dev.from_device(B, kind)


_stream_rng = None      # the reference keeps drawing from numpy's global generator between the streaming calls


@cr('math.init_qr_streaming')
def init_qr_streaming(Ai, r, q, seed=None):
    """First block of the streaming randomized QR (pyLOM/vmmath/svd.py:176-200): randomized_qr that also returns
    the sketch Yi (m_i, r).  Seeds the generator the following update_qr_streaming calls draw from."""
    global _stream_rng
    Ad, kind = _dev.to_device(Ai, "Ai")
    _stream_rng = _np.random.RandomState(int(time.time()) if seed is None else int(seed))
    Qi, B, Yi = _randomized_qr_dev(Ad, r, q, None, rng=_stream_rng)
    return _dev.from_device(Qi, kind), _dev.from_device(B, kind), _dev.from_device(Yi, kind)


@cr('math.qr_iteration')
def update_qr_streaming(Ai, Q1, B1, Yo, r, q):
    """Next block of snapshots -- same rows, new columns (pyLOM/vmmath/svd.py:202-226): Ai (m_i, n2), Q1 (m_i, r),
    B1 (r, n1), Yo (m_i, r)  ->  Q2 (m_i, r), B2 (r, n1 + n2), Yo + Yn."""
    global _stream_rng
    eng = _engine
    Ad, kind = _dev.to_device(Ai, "Ai")
    Q1d, _ = _dev.to_device(Q1, "Q1")
    B1d, _ = _dev.to_device(B1, "B1")
    Yod, _ = _dev.to_device(Yo, "Yo")
    m, n = Ad.shape
    r = int(r)
    if m < r:       # the reference has no r <= n2 restriction here: omega is (n2, r), Yn is (m, r) (svd.py:202-226)
        raise ValueError(f"every rank needs at least r rows (got m_i={m} < r={r})")
    if _stream_rng is None:
        _stream_rng = _np.random.RandomState()
    Yn = _power_sketch(Ad, _sketch_matrix(n, r, _stream_rng, Ad.device), q, eng)
    Ynew = Yod + Yn
    Q2, _R = _tsqr_dev(Ynew, eng)
    Q2Q1 = parall.mpi_reduce(eng.matmul_tn(Q2, Q1d), op='sum', all=True)        # (r, r)
    B2n = parall.mpi_reduce(eng.matmul_tn(Q2, Ad), op='sum', all=True)          # (r, n2)
    B2 = torch.cat((eng.matmul(Q2Q1, B1d), B2n), dim=1)
    return _dev.from_device(Q2, kind), _dev.from_device(B2, kind), _dev.from_device(Ynew, kind)


@cr('math.randomized_svd')
def randomized_svd(Ai, r, q, seed=-1):
    """Randomized SVD (pyLOM/vmmath/svd.py:254-273; drandomized_svd src/svd.c:1453-1519):
    Ai (m_i, n) -> Ui (m_i, r), S (r), V (r, n) = V^T."""
    Ad, kind = _dev.to_device(Ai, "Ai")
    Ui, S, V = _randomized_svd_dev(Ad, r, q, seed)
    return _dev.from_device(Ui, kind), _dev.from_device(S, kind), _dev.from_device(V, kind)
