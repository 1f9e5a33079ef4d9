"""Executable specification (numpy) of the fused small-n TSQR kernels in pyloworder_b200/csrc/tsqr_small.cu.

Same decomposition, same order of operations, same stored quantities as the CUDA code (n <= NP, NP = 32 or 64):

  * rows are cut into STRIPS; a strip = a dense HEAD of NP rows followed by TALL TILES of TB = 512 rows (the last tile
    of the last strip may be partial).  Strips are independent chains; their NP x NP triangles are stacked and
    reduced by the generic CAQR path;
  * the head gets an ordinary (unit-lower) Householder QR with a full NP x NP compact-WY T (dlarft recurrence);
  * every tile gets the structured QR of [R; tile] with reflectors [e_j; v_j] (R = the strip's running triangle),
    LEFT-LOOKING in sub-panels of 8 columns: the sub-panel is loaded, the tile's earlier reflectors are applied as one
    block reflector with their accumulated T (W = R[prev, J] + V_prev^T C, W' = T_prev^T W, R[prev, J] -= W',
    C -= V_prev W'), then 8 structured Householder column steps (reflector columns stay unscaled during the chain,
    the 8 x 8 T8 comes from the chain's own inner products), and T grows by 8 columns:
    T[prev, J] = -T_prev (G T8) with the Gram block G = V_prev^T V_J;
  * stored per tile: V (in place of the tile) and T;
  * pass 2 (apply Q to [C; 0], tiles in reverse order):  X = T C,  C -= X,  U_tile = -V X  (the zero block of
    the target makes this ONE K = NP GEMM per tile);  head:  U_head = C - Y T (Y^T C).

tests/test_model.py checks this model against numpy.linalg.qr; tests/test_gpu_parity.py checks the CUDA kernels
against the same properties and against the oracle.
"""
import numpy as np

TB = 512
SPW = 8


def house(alpha, sigma2):
    """beta, tau, scale of the reflector that maps [alpha; x] (x^T x = sigma2) to [beta; 0]."""
    if sigma2 == 0.0:
        return alpha, 0.0, 0.0
    nrm = np.sqrt(alpha * alpha + sigma2)
    beta = -np.copysign(nrm, alpha)
    return beta, (beta - alpha) / beta, 1.0 / (alpha - beta)


def plan(m, NP, nstrips_target=296, min_tiles=8):
    """Strips (r0, ntiles, rows of the last tile).  Every strip = NP head rows + whole tiles; the partial tile (if any)
    goes to the last strip.  Same arithmetic as small_plan() in tsqr_small.cu."""
    ns = max(1, min(nstrips_target, m // (NP + min_tiles * TB)))
    body = m - ns * NP
    Tt, part = body // TB, body % TB
    q, rem = Tt // ns, Tt % ns
    strips = []
    for i in range(ns):
        r0 = i * NP + TB * (i * q + min(i, rem))
        nt = q + (1 if i < rem else 0)
        last = TB
        if i == ns - 1 and part:
            nt, last = nt + 1, part
        strips.append((r0, nt, last))
    return strips


def chain(C, R, c0):
    """8 structured Householder steps on [R[:, J]; C] (C = the m x 8 sub-panel, J = c0..c0+7).  Returns T8."""
    T8 = np.zeros((SPW, SPW))
    sc = np.zeros(SPW)
    for jj in range(SPW):
        j = c0 + jj
        x = C[:, jj].copy()
        tot = x @ C                                  # dots with all 8 columns (earlier ones are unscaled reflectors)
        beta, tau, scale = house(R[j, j], tot[jj])
        for c in range(jj + 1, SPW):
            w = tau * (R[j, c0 + c] + scale * tot[c])
            R[j, c0 + c] -= w
            C[:, c] -= x * (scale * w)
        R[j, j] = beta
        sc[jj] = scale
        z = sc[:jj] * scale * tot[:jj]               # v_l^T v_j, l < jj
        T8[:jj, jj] = -tau * (T8[:jj, :jj] @ z)
        T8[jj, jj] = tau
    C *= sc                                          # scale the reflector columns once at the end
    return T8


def factor_tile(At, R, NP):
    """Left-looking structured QR of [R; At]; At is overwritten by V.  Returns T (NP x NP upper)."""
    T = np.zeros((NP, NP))
    for k in range(NP // SPW):
        c0 = SPW * k
        J, P = slice(c0, c0 + SPW), slice(0, c0)
        C = At[:, J]                                 # view
        if k:
            W = R[P, J] + At[:, P].T @ C
            Wp = T[P, P].T @ W
            R[P, J] -= Wp
            C -= At[:, P] @ Wp
        T8 = chain(C, R, c0)
        T[J, J] = T8
        if k:
            G = At[:, P].T @ C
            T[P, J] = -T[P, P] @ (G @ T8)
    return T


def factor_head(H, NP):
    """Dense Householder QR of the NP x NP head, in place (R above, unit-lower Y below).  Returns T (NP x NP)."""
    T = np.zeros((NP, NP))
    for j in range(NP):
        x = H[j + 1:, j].copy()
        beta, tau, scale = house(H[j, j], x @ x)
        v = x * scale
        w = tau * (H[j, j + 1:] + v @ H[j + 1:, j + 1:])
        H[j, j + 1:] -= w
        H[j + 1:, j + 1:] -= np.outer(v, w)
        H[j, j] = beta
        H[j + 1:, j] = v
        Y = np.tril(H[:, :j], -1) + np.eye(NP)[:, :j]
        yj = np.zeros(NP); yj[j] = 1.0; yj[j + 1:] = v
        T[:j, j] = -tau * (T[:j, :j] @ (Y.T @ yj))
        T[j, j] = tau
    return T


def factor(A, NP=64, **kw):
    """Pass 1 on an m x n matrix (n <= NP).  Returns the state pass 2 needs and the stacked strip triangles."""
    m, n = A.shape
    strips = plan(m, NP, **kw)
    V = np.zeros((m, NP))
    V[:, :n] = A
    Th, Tt = [], []
    Rstack = np.zeros((len(strips) * NP, NP))
    for i, (r0, nt, last) in enumerate(strips):
        head = V[r0:r0 + NP]                                   # view: factored in place
        Th.append(factor_head(head, NP))
        R = np.triu(head).copy()
        Ts = []
        for t in range(nt):
            t0 = r0 + NP + t * TB
            rows = last if t == nt - 1 else TB
            Ts.append(factor_tile(V[t0:t0 + rows], R, NP))
        Tt.append(Ts)
        Rstack[i * NP:(i + 1) * NP] = np.triu(R)
    return {"V": V, "Th": Th, "Tt": Tt, "strips": strips, "NP": NP, "n": n}, Rstack


def apply_q(state, Bstack):
    """Pass 2: U = Q [B_s per strip].  Bstack ((nstrips NP) x k)."""
    V, NP = state["V"], state["NP"]
    U = np.zeros((V.shape[0], Bstack.shape[1]))
    for i, (r0, nt, last) in enumerate(state["strips"]):
        C = Bstack[i * NP:(i + 1) * NP].copy()
        for t in range(nt - 1, -1, -1):
            t0 = r0 + NP + t * TB
            rows = last if t == nt - 1 else TB
            X = state["Tt"][i][t] @ C
            C -= X
            U[t0:t0 + rows] = -V[t0:t0 + rows] @ X
        Y = np.tril(V[r0:r0 + NP], -1) + np.eye(NP)
        U[r0:r0 + NP] = C - Y @ (state["Th"][i] @ (Y.T @ C))
    return U


def qr(A, NP=64, **kw):
    """Thin QR through the two passes with a LAPACK QR of the strip triangles in between (the CUDA path uses the
    generic CAQR kernels for that level)."""
    n = A.shape[1]
    state, Rstack = factor(A, NP, **kw)
    Q2, R = np.linalg.qr(Rstack[:, :n])     # (nstrips NP) x n
    Q = apply_q(state, Q2)
    return Q, R
