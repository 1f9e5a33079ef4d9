"""CPU oracle for the POD / TSQR-SVD hot path  --  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the reference algorithm (pyLOM 3.2.8).  It is the
*checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``pyloworder_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the reference's own unmodified
``pyLOM/vmmath/{maths,averaging,truncation,svd,stats}.py`` and ``pyLOM/POD/wrapper.py``
(loaded under a stub ``pyLOM.utils``; P simulated ranks) and stores inputs/outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks every function here against those
fixtures, and (when built) against ``oracle/_ref/libpylom_ref.so`` compiled from the
reference's C sources.

Third-party arithmetic: the reference delegates QR/SVD/GEMM to LAPACK/BLAS
(OpenBLAS 0.3.17 or oneMKL 2024.2.0.634, ``options.cfg:45-46``) or numpy; here it is
numpy's bundled OpenBLAS.  LAPACK results are unique only up to the sign of each mode
(and rotation inside degenerate clusters), so all comparisons are sign-invariant.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# partitioning
# --------------------------------------------------------------------------------------
def worksplit(istart: int, iend: int, whoAmI: int, nWorkers: int):
    """Contiguous row range of worker ``whoAmI``  (pyLOM/utils/parall.py:24-48)."""
    istart_l, iend_l = istart, iend
    irange = iend - istart
    if nWorkers < irange:
        rangePerProcess = int(np.floor(irange / nWorkers))
        istart_l = istart + whoAmI * rangePerProcess
        iend_l = istart_l + rangePerProcess
        remainder = irange - rangePerProcess * nWorkers
        if remainder > whoAmI:
            istart_l += whoAmI
            iend_l += whoAmI + 1
        else:
            istart_l += remainder
            iend_l += remainder
    else:
        istart_l = whoAmI if whoAmI < iend else iend
        iend_l = whoAmI + 1 if whoAmI < iend else iend
    return istart_l, iend_l


# --------------------------------------------------------------------------------------
# averaging
# --------------------------------------------------------------------------------------
def temporal_mean(X: np.ndarray) -> np.ndarray:
    """Row mean over the n snapshots (pyLOM/vmmath/averaging.py:17-29, src/averaging.c:29-46)."""
    return np.mean(X, axis=1)


def subtract_mean(X: np.ndarray, X_mean: np.ndarray) -> np.ndarray:
    """out[i,j] = X[i,j] - X_mean[i] (pyLOM/vmmath/averaging.py:31-44, src/averaging.c:109-124)."""
    return X - X_mean[:, None]


def temporal_variance(X: np.ndarray, X_mean: np.ndarray) -> np.ndarray:
    """Population variance per row, keepdims (pyLOM/vmmath/averaging.py:46-59, src/averaging.c:70-90)."""
    return np.var(X, axis=1, keepdims=True)


def norm_variance(X, X_mean, X_var):
    """(X - X_mean)/X_var (pyLOM/vmmath/averaging.py:61-74)."""
    return subtract_mean(X, X_mean) / X_var


# --------------------------------------------------------------------------------------
# small dense algebra
# --------------------------------------------------------------------------------------
def matmul(A, B):
    """C = A x B (pyLOM/vmmath/maths.py:76-90, src/vector_matrix.c:234-242)."""
    return np.matmul(A, B)


def vecmat(v, A):
    """C[i,:] = v[i]*A[i,:] (pyLOM/vmmath/maths.py:112-129, src/vector_matrix.c:401-414)."""
    return v[:, None] * A


def qr(A):
    """Thin Householder QR (pyLOM/vmmath/svd.py:28-36; geqrf+orgqr in src/svd.c:280-321)."""
    return np.linalg.qr(A)


def svd(A):
    """Thin SVD, S descending, V returned as V^T (pyLOM/vmmath/svd.py:38-47, src/svd.c:83-139)."""
    return np.linalg.svd(A, full_matrices=False)


def next_power_of_2(n: int) -> int:
    """pyLOM/vmmath/svd.py:17-25, src/svd.c:409-414."""
    p = 1
    if n and not (n & (n - 1)):
        return n
    while p < n:
        p <<= 1
    return p


# --------------------------------------------------------------------------------------
# TSQR (butterfly, P simulated ranks in one process)
# --------------------------------------------------------------------------------------
def tsqr(A_list):
    """Parallel QR of a row-partitioned matrix, P = len(A_list) ranks.

    Follows pyLOM/vmmath/svd.py:49-118 (src/svd.c:565-676) step by step; every rank's
    control flow is executed level-synchronously, a send/recv pair becomes a mailbox
    hand-off.  Returns ([Q_i], [R_i]) -- R is identical on every rank.
    """
    P = len(A_list)
    n = A_list[0].shape[1]
    dt = A_list[0].dtype
    nextPower = next_power_of_2(P)
    nlevels = int(np.log2(nextPower))
    Q1 = [None] * P
    R = [None] * P
    for r in range(P):
        Q1[r], R[r] = qr(A_list[r])                                   # svd.py:60
    QW = [np.eye(n, dtype=dt) for _ in range(P)]                      # svd.py:63
    C = [np.zeros((2 * n, n), dt) for _ in range(P)]
    Q2l = [np.zeros((2 * n * nlevels, n), dt) for _ in range(P)]
    blevel = 1
    for ilevel in range(nlevels):                                     # svd.py:67-84
        mailbox = {}
        for r in range(P):
            C[r][:n, :] = R[r]
            prank = r ^ blevel
            if r & blevel and prank < P:
                mailbox[prank] = R[r].copy()                          # mpi_send(R, prank)
        for r in range(P):
            prank = r ^ blevel
            if not (r & blevel) and prank < P:
                R[r] = mailbox[r]                                     # mpi_recv
                C[r][n:, :] = R[r]
                Q2i, R[r] = qr(C[r])
                Q2l[r][2 * n * ilevel:2 * n * ilevel + 2 * n, :] = Q2i
        blevel <<= 1
    if P > 1:                                                         # svd.py:87-115
        blevel = 1 << (nlevels - 1)
        mask = blevel - 1
    for ilevel in reversed(range(nlevels)):
        mailbox = {}
        Q2i_of = {}
        for r in range(P):                                            # senders first
            if r & mask == 0:
                Cb = Q2l[r][2 * n * ilevel:2 * n * ilevel + 2 * n, :]
                Q2i = matmul(Cb, QW[r])
                Q2i_of[r] = Q2i
                prank = r ^ blevel
                if not (r & blevel) and prank < P:
                    msg = np.empty((2 * n, n), dt)
                    msg[:n, :] = R[r]
                    msg[n:, :] = Q2i[n:, :]
                    QW[r] = Q2i[:n, :].copy()
                    mailbox[prank] = msg
        for r in range(P):                                            # then receivers
            if r & mask == 0:
                prank = r ^ blevel
                if r & blevel and prank < P:
                    msg = mailbox[r]
                    R[r] = msg[:n, :].copy()
                    QW[r] = msg[n:, :].copy()
        blevel >>= 1
        mask >>= 1
    Q = [matmul(Q1[r], QW[r]) for r in range(P)]                      # svd.py:117
    return Q, R


def tsqr_svd(A_list):
    """TSQR-based SVD on P simulated ranks (pyLOM/vmmath/svd.py:227-252, src/svd.c:678-712).

    Accepts a single 2-D array (one rank) or a list of row blocks.  Returns
    ([U_i], S, V) with V = V^T (n x n).
    """
    single = isinstance(A_list, np.ndarray)
    if single:
        A_list = [A_list]
    Q, R = tsqr(A_list)
    Ur, S, V = svd(R[0])
    U = [matmul(Qi, Ur) for Qi in Q]
    return (U[0] if single else U), S, V


def matmulp(A_list, B_list):
    """C = sum over ranks of A_i x B_i (pyLOM/vmmath/maths.py:93-110; dmatmulp, src/vector_matrix.c:344-356:
    local GEMM + MPI_Allreduce).  A_i (M, q_i), B_i (q_i, N): the inner dimension is the distributed one."""
    tot = np.matmul(A_list[0], B_list[0])
    for A, B in zip(A_list[1:], B_list[1:]):
        tot = tot + np.matmul(A, B)
    return tot


def sketch_matrix(n: int, r: int, seed: int) -> np.ndarray:
    """The Gaussian-free sketch of randomized_qr: ``np.random.seed(seed); np.random.rand(n, r)``
    (pyLOM/vmmath/svd.py:131-133).  RandomState(seed) is the same MT19937 stream without touching the global state."""
    return np.random.RandomState(seed).rand(n, r)


def randomized_qr(A_list, r: int, q: int, seed: int):
    """Randomized range finder with q power iterations on P simulated ranks (pyLOM/vmmath/svd.py:120-144;
    drandomized_qr, src/svd.c:1267-1319).  Returns ([Q_i (m_i, r)], B (r, n))."""
    single = isinstance(A_list, np.ndarray)
    As = [A_list] if single else A_list
    n = As[0].shape[1]
    omega = sketch_matrix(n, r, seed)
    Y = [matmul(A, omega) for A in As]
    for _ in range(q):
        Q, _R = tsqr(Y)
        Q2 = matmulp([A.T for A in As], Q)
        Y = [matmul(A, Q2) for A in As]
    Q, _R = tsqr(Y)
    B = matmulp([Qi.T for Qi in Q], As)
    return (Q[0] if single else Q), B


_stream_rng = None      # the reference keeps drawing from numpy's global generator between the streaming calls


def init_qr_streaming(A_list, r: int, q: int, seed: int):
    """First block of the streaming randomized QR (pyLOM/vmmath/svd.py:176-200): randomized_qr that also returns the
    sketch Y_i.  Seeds the generator the following update_qr_streaming calls keep drawing from."""
    global _stream_rng
    single = isinstance(A_list, np.ndarray)
    As = [A_list] if single else A_list
    n = As[0].shape[1]
    _stream_rng = np.random.RandomState(seed)
    omega = _stream_rng.rand(n, r)
    Y = [matmul(A, omega) for A in As]
    for _ in range(q):
        Q, _R = tsqr(Y)
        Q2 = matmulp([A.T for A in As], Q)
        Y = [matmul(A, Q2) for A in As]
    Q, _R = tsqr(Y)
    B = matmulp([Qi.T for Qi in Q], As)
    return (Q[0], B, Y[0]) if single else (Q, B, Y)


def update_qr_streaming(A_list, Q1, B1, Yo, r: int, q: int):
    """Next block of snapshots (same rows, new columns) (pyLOM/vmmath/svd.py:202-226): a fresh sketch of the new
    block is power-iterated, added to the running Y, re-orthonormalised; B is carried over through Q2^T Q1."""
    single = isinstance(A_list, np.ndarray)
    As = [A_list] if single else A_list
    Q1s = [Q1] if single else Q1
    Yos = [Yo] if single else Yo
    n = As[0].shape[1]
    omega = _stream_rng.rand(n, r)
    Yn = [matmul(A, omega) for A in As]
    for _ in range(q):
        Qp, _R = tsqr(Yn)
        O2 = matmulp([A.T for A in As], Qp)
        Yn = [matmul(A, O2) for A in As]
    Yos = [a + b for a, b in zip(Yos, Yn)]
    Q2, _R = tsqr(Yos)
    Q2Q1 = matmulp([x.T for x in Q2], Q1s)
    B2 = np.hstack((matmul(Q2Q1, B1), matmulp([x.T for x in Q2], As)))
    return (Q2[0], B2, Yos[0]) if single else (Q2, B2, Yos)


def randomized_svd(A_list, r: int, q: int, seed: int):
    """Randomized SVD (pyLOM/vmmath/svd.py:254-273; drandomized_svd, src/svd.c:1453-1519):
    ([U_i (m_i, r)], S (r), V (r, n))."""
    single = isinstance(A_list, np.ndarray)
    As = [A_list] if single else A_list
    Q, B = randomized_qr(As, r, q, seed)
    Ur, S, V = svd(B)
    U = [matmul(Qi, Ur) for Qi in Q]
    return (U[0] if single else U), S, V


# --------------------------------------------------------------------------------------
# POD
# --------------------------------------------------------------------------------------
def pod_run(X_list, remove_mean: bool = True, divide_variance: bool = False, randomized: bool = False, r: int = 1,
            q: int = 3, seed: int = -1):
    """POD.run on P simulated ranks (pyLOM/POD/wrapper.py:16-51).  ``X_list`` may be a
    single array."""
    single = isinstance(X_list, np.ndarray)
    Xs = [X_list] if single else X_list
    if remove_mean and divide_variance:
        Ys = [norm_variance(X, temporal_mean(X), temporal_variance(X, temporal_mean(X))) for X in Xs]
    elif remove_mean:
        Ys = [subtract_mean(X, temporal_mean(X)) for X in Xs]
    else:
        Ys = [X.copy() for X in Xs]
    U, S, V = randomized_svd(Ys, r, q, seed) if randomized else tsqr_svd(Ys)
    return (U[0] if single else U), S, V


# --------------------------------------------------------------------------------------
# DMD on the POD basis
# --------------------------------------------------------------------------------------
def vandermonde(real, imag, m, n):
    """Vand[:, k] = (real + i imag)**k, k < n  (pyLOM/vmmath/maths.py:202-222)."""
    Vand = np.zeros((m, n), dtype=np.complex128)
    for icol in range(n):
        Vand[:, icol] = (real + imag * 1j) ** icol
    return Vand


def vandermonde_time(real, imag, m, time):
    """Vand[:, it] = (real + i imag)**t  (pyLOM/vmmath/maths.py:224-246)."""
    Vand = np.zeros((m, time.shape[0]), dtype=np.complex128)
    for it, t in enumerate(time):
        Vand[:, it] = (real + imag * 1j) ** t
    return Vand


def dmd_order_modes(muReal, muImag, Phi, bJov):
    """Sort by decreasing |b| and put the positive-imaginary member of every conjugate pair first
    (pyLOM/DMD/wrapper.py:18-47, statement for statement -- including what it does to Phi.imag and bJov.imag)."""
    order = np.flip(np.abs(bJov).argsort())
    muReal = muReal[order]
    muImag = muImag[order]
    Phi = np.transpose(np.transpose(Phi)[order])
    bJov = bJov[order]
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
            Phi.imag[:, ii] = Phi.imag[:, ii + 1]
            Phi.imag[:, ii + 1] = -Phi.imag[:, ii + 1]
            p = True
            continue
        if iimag > 0:
            p = True
            continue
    return muReal, muImag, Phi, bJov


def dmd_run(X_list, r, remove_mean: bool = True):
    """DMD.run on P simulated ranks (pyLOM/DMD/wrapper.py:50-117): SVD of the first n-1 snapshots, projected linear
    map Atilde = U^T Y2 V S^-1, its eigen-decomposition, modes Phi = Y2 V S^-1 w / mu, Jovanovic amplitudes."""
    single = isinstance(X_list, np.ndarray)
    Xs = [X_list] if single else X_list
    Ys = [subtract_mean(X, temporal_mean(X)) for X in Xs] if remove_mean else [X.copy() for X in Xs]
    U, S, VT = tsqr_svd([np.ascontiguousarray(Y[:, :-1]) for Y in Ys])
    N = int(r) if r >= 1 else compute_truncation_residual(S, r)
    U = [Ui[:, :N] for Ui in U]; S = S[:N]; VT = VT[:N, :]
    aux1 = matmulp([Ui.T for Ui in U], [Y[:, 1:] for Y in Ys])
    aux2 = np.transpose(vecmat(1. / S, VT))
    Atilde = matmul(aux1, aux2)
    mu, w = np.linalg.eig(Atilde)
    muReal, muImag = np.real(mu), np.imag(mu)
    Phi = [matmul(matmul(matmul(Y[:, 1:], np.transpose(VT)), np.diag(1 / S)), w) / (muReal + muImag * 1J) for Y in Ys]
    Vand = vandermonde(muReal, muImag, muReal.shape[0], Ys[0].shape[1] - 1)
    P = matmul(np.transpose(np.conj(w)), w) * np.conj(matmul(Vand, np.transpose(np.conj(Vand))))
    Pl = np.linalg.cholesky(P)
    G = matmul(np.diag(S), VT)
    q = np.conj(np.diag(matmul(matmul(Vand, np.transpose(np.conj(G))), w)))
    bJov = matmul(np.linalg.inv(np.transpose(np.conj(Pl))), matmul(np.linalg.inv(Pl), q))
    Phi_all = np.vstack(Phi)
    muReal, muImag, Phi_all, bJov = dmd_order_modes(muReal, muImag, Phi_all, bJov)
    if single:
        return muReal, muImag, Phi_all, bJov
    cuts = np.cumsum([0] + [Y.shape[0] for Y in Ys])
    return muReal, muImag, [Phi_all[cuts[k]:cuts[k + 1]] for k in range(len(Ys))], bJov


def dmd_frequency_damping(real, imag, dt):
    """pyLOM/DMD/wrapper.py:119-130."""
    mod = np.sqrt(real * real + imag * imag)
    arg = np.arctan2(imag, real)
    return np.log(mod) / dt, arg / dt


def dmd_mode_computation(X, V, S, W):
    """pyLOM/DMD/wrapper.py:132-138."""
    return matmul(matmul(matmul(X, np.transpose(V)), np.diag(1 / S)), np.abs(W))


def dmd_reconstruction_jovanovic(Phi, real, imag, t, bJov):
    """pyLOM/DMD/wrapper.py:140-146."""
    Vand = vandermonde_time(real, imag, real.shape[0], t)
    return matmul(Phi, matmul(np.diag(bJov), Vand)).real


def vector_norm(v, start=0):
    """pyLOM/vmmath/maths.py:47-59."""
    return np.linalg.norm(v[start:], 2)


def vector_sum(v, start=0):
    """pyLOM/vmmath/maths.py:32-44."""
    return np.sum(v[start:])


def compute_truncation_residual(S, r):
    """Number of modes to keep (pyLOM/vmmath/truncation.py:17-39, src/truncation.c:46-74)."""
    N = 0
    if r > 0:
        normS = vector_norm(S, 0)
        for ii in range(S.shape[0]):
            accumulative = vector_norm(S, ii) / normS
            if accumulative < r:
                break
            N += 1
    else:
        r = abs(r)
        normS = vector_sum(S, 0)
        accumulative = 0
        for ii in range(S.shape[0]):
            accumulative += S[ii] / normS
            N += 1
            if accumulative > r:
                break
    return N


def truncate(U, S, V, r=1e-8):
    """pyLOM/POD/wrapper.py:55-82."""
    N = int(r) if r >= 1 else compute_truncation_residual(S, r)
    return U[:, :N], S[:N], V[:N, :]


def reconstruct(U, S, V):
    """X = U diag(S) V (pyLOM/POD/wrapper.py:86-103); the mean is NOT re-added."""
    return matmul(U, vecmat(S, V))


def RMSE(A_list, B_list, relative: bool = True):
    """sqrt(sum((A-B)^2)/sum(A^2)) with both sums reduced over ranks (pyLOM/vmmath/stats.py:17-34)."""
    if isinstance(A_list, np.ndarray):
        A_list, B_list = [A_list], [B_list]
    s1 = sum(float(np.sum((A - B) * (A - B))) for A, B in zip(A_list, B_list))
    if relative:
        s2 = sum(float(np.sum(A * A)) for A in A_list)
    else:
        s2 = float(np.prod(np.sum([np.array(A.shape) for A in A_list], axis=0)))
    return np.sqrt(s1 / s2)


# --------------------------------------------------------------------------------------
# sign-/rotation-invariant comparators used by the parity tests
# --------------------------------------------------------------------------------------
def energy(A_list, B_list):
    """Reconstruction energy 1 - sum((A-B)^2)/sum(A^2) over all ranks (pyLOM/vmmath/truncation.py:40-60)."""
    if isinstance(A_list, np.ndarray):
        A_list, B_list = [A_list], [B_list]
    num = sum(np.sum((A - B) ** 2) for A, B in zip(A_list, B_list))
    den = sum(np.sum(A ** 2) for A in A_list)
    return 1 - num / den


def extract_modes(U, ivar, npoints, modes=(), reshape=True):
    """pyLOM/POD/utils.py:19-42."""
    nvars = U.shape[0] // npoints
    if len(modes) == 0:
        modes = np.arange(1, U.shape[1] + 1, dtype=np.int32)
    out = np.zeros((npoints, len(modes)), U.dtype)
    for i, m in enumerate(modes):
        out[:, i] = U[ivar - 1:nvars * npoints:nvars, m - 1]
    return out.reshape((len(modes) * npoints,), order='C') if reshape else out


def compare_svd(U_ref, S_ref, V_ref, U, S, V, gap_tol=1e-6, floor=1e-8):
    """Return a dict of parity metrics (SURVEY.md section 8d acceptance).

    * ``sigma_rel``  max_k |s_k - s_k_ref| / s_1_ref
    * ``sigma_rel_each`` max over k with s_k/s_1 >= 1e-6 of |s_k - s_k_ref|/s_k_ref
    * ``mode_min``   min over well separated modes of |<u_ref_k, u_k>|
    * ``vmode_min``  same for rows of V
    """
    S_ref = np.asarray(S_ref); S = np.asarray(S)
    s1 = S_ref[0]
    out = {"sigma_rel": float(np.max(np.abs(S - S_ref)) / s1)}
    big = S_ref / s1 >= 1e-6
    out["sigma_rel_each"] = float(np.max(np.abs(S[big] - S_ref[big]) / S_ref[big])) if big.any() else 0.0
    n = S_ref.shape[0]
    sep = np.zeros(n, bool)
    for k in range(n):
        lo = S_ref[k - 1] - S_ref[k] if k > 0 else np.inf
        hi = S_ref[k] - S_ref[k + 1] if k < n - 1 else np.inf
        sep[k] = (min(lo, hi) / s1 >= gap_tol) and (S_ref[k] / s1 >= floor)
    out["n_separated"] = int(sep.sum())
    if sep.any():
        du = np.abs(np.einsum("ik,ik->k", U_ref[:, sep], U[:, sep]))
        dv = np.abs(np.einsum("ki,ki->k", V_ref[sep, :], V[sep, :]))
        out["mode_min"] = float(du.min())
        out["vmode_min"] = float(dv.min())
    else:
        out["mode_min"] = out["vmode_min"] = 1.0
    return out
