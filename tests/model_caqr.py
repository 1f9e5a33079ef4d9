"""Executable numpy model of the device algorithm -- host-side spec.

It mirrors the CUDA decomposition one-to-one (same panel / level / strip / tile loops, same
buffers: padded work matrix ``Vb``, per-tile ``T`` store, upper-level ``Vup`` store, panel
scratch ``Ptmp``) so that indexing and algebra can be validated on the CPU, where the kernels
cannot run.  ``tests/test_model.py`` checks it against LAPACK.  Not used by the product.
"""
import numpy as np

NB = 32      # panel width
G = 4        # NB-row blocks per tile  (tile = G*NB = 128 rows)


def plan_levels(m_act, strips_target=592, smax=8):
    """Level geometry for one panel: list of dicts(nblk, bs, ntiles, s, nstrips)."""
    levels = []
    nblk, bs = -(-m_act // NB), NB
    while True:
        ntiles = -(-nblk // G)
        if ntiles > 148:
            s = min(smax, max(1, ntiles // strips_target))
        else:
            s = min(ntiles, 4)
        nstrips = -(-ntiles // s)
        levels.append(dict(nblk=nblk, bs=bs, ntiles=ntiles, s=s, nstrips=nstrips))
        if nstrips == 1:
            return levels
        nblk, bs = nstrips, bs * G * s


def _block_rows(row0, lv, k):
    r = row0 + k * lv["bs"]
    return slice(r, r + NB)


def _house_stack(Rp, body):
    """Householder QR of the stacked [Rp (NB x NB, rows<j inactive for column j); body].

    Returns (Vp unit-lower NBxNB explicit, Vbody, T, R upper)."""
    Rp = Rp.copy(); body = body.copy()
    nb = Rp.shape[1]
    tau = np.zeros(nb)
    for j in range(nb):
        alpha = Rp[j, j]
        xp = Rp[j + 1:, j]
        xb = body[:, j]
        sig = xp @ xp + xb @ xb
        if sig == 0.0:
            tau[j] = 0.0
            Rp[j + 1:, j] = 0.0
            body[:, j] = 0.0
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + sig), alpha)
        tau[j] = (beta - alpha) / beta
        sc = 1.0 / (alpha - beta)
        vp = xp * sc
        vb = xb * sc
        # trailing columns inside the panel
        w = tau[j] * (Rp[j, j + 1:] + vp @ Rp[j + 1:, j + 1:] + vb @ body[:, j + 1:])
        Rp[j, j + 1:] -= w
        Rp[j + 1:, j + 1:] -= np.outer(vp, w)
        body[:, j + 1:] -= np.outer(vb, w)
        Rp[j, j] = beta
        Rp[j + 1:, j] = vp
        body[:, j] = vb
    Vp = np.tril(Rp, -1) + np.eye(nb)
    R = np.triu(Rp)
    # T from the Gram matrix:  T^{-1} = striu(V^T V) + diag(1/tau)
    Gm = Vp.T @ Vp + body.T @ body
    Tinv = np.triu(Gm, 1)
    T = np.zeros((nb, nb))
    for j in range(nb):
        if tau[j] != 0.0:
            T[j, j] = tau[j]
            T[:j, j] = -tau[j] * (T[:j, :j] @ Tinv[:j, j])
    return Vp, body, T, R


class CAQR:
    """Factor A (m x n) in a padded buffer; keep what the device keeps."""

    def __init__(self, A):
        m, n = A.shape
        self.m, self.n = m, n
        self.npad = -(-n // NB) * NB
        self.K = self.npad // NB
        self.Mrows = m + self.npad + NB
        self.Vb = np.zeros((self.Mrows, self.npad))
        self.Vb[:m, :n] = A
        self.plans = []
        self.T = {}      # (p, level, tile) -> NBxNB
        self.Vup = {}    # (p, level, tile) -> (G*NB) x NB explicit V, levels >= 1 (0-based)

    # ---- tile helpers ---------------------------------------------------------------------
    def _tile_blocks(self, lv, t):
        return [k for k in range(t * G, t * G + G) if k < lv["nblk"]]

    def _load_panel_block(self, row0, col0, lv, li, k):
        blk = self.Vb[_block_rows(row0, lv, k), col0:col0 + NB].copy()
        if li > 0:
            blk = np.triu(blk)          # lower part holds level-0 reflectors: mask
        return blk

    # ---- factorisation --------------------------------------------------------------------
    def factor(self):
        for p in range(self.K):
            row0 = col0 = p * NB
            levels = plan_levels(self.m - row0)
            self.plans.append(levels)
            for li, lv in enumerate(levels):
                self._panel_level(p, row0, col0, li, lv)
                if col0 + NB < self.npad:
                    self._update_level(p, row0, li, lv, self.Vb, col0 + NB, self.npad, forward=True)
        self.R = np.triu(self.Vb[:self.n, :self.n]).copy()
        return self.R

    def _panel_level(self, p, row0, col0, li, lv):
        for j in range(lv["nstrips"]):
            t0 = j * lv["s"]
            piv = t0 * G
            Rp = self._load_panel_block(row0, col0, lv, li, piv)
            for i in range(lv["s"]):
                t = t0 + i
                if t >= lv["ntiles"]:
                    break
                blocks = self._tile_blocks(lv, t)
                body_blocks = blocks[1:] if i == 0 else blocks
                body = np.vstack([self._load_panel_block(row0, col0, lv, li, k) for k in body_blocks]) \
                    if body_blocks else np.zeros((0, NB))
                if i > 0:
                    Rp = np.triu(Rp)
                Vp, Vbody, T, R = _house_stack(Rp, body)
                self.T[(p, li, t)] = T
                # explicit tile V (G*NB rows, missing blocks zero)
                Vt = np.zeros((G * NB, NB))
                off = 0
                if i == 0:
                    Vt[:NB] = Vp
                    off = 1
                for q, k in enumerate(body_blocks):
                    Vt[(q + off) * NB:(q + off + 1) * NB] = Vbody[q * NB:(q + 1) * NB]
                if li == 0:
                    for q, k in enumerate(body_blocks):
                        self.Vb[_block_rows(row0, lv, k), col0:col0 + NB] = Vbody[q * NB:(q + 1) * NB]
                    if i == 0:
                        self.Vb[_block_rows(row0, lv, piv), col0:col0 + NB] = np.tril(Vp, -1)
                else:
                    self.Vup[(p, li, t)] = Vt
                Rp = R
            # R of the strip goes to the upper triangle of its pivot block
            sl = _block_rows(row0, lv, piv)
            low = np.tril(self.Vb[sl, col0:col0 + NB], -1)
            self.Vb[sl, col0:col0 + NB] = low + np.triu(Rp)

    def _tile_V(self, p, row0, col0, li, lv, t, first):
        """Explicit (G*NB) x NB reflector block of tile t as the update kernel sees it."""
        if li > 0:
            return self.Vup[(p, li, t)]
        Vt = np.zeros((G * NB, NB))
        for q, k in enumerate(self._tile_blocks(lv, t)):
            Vt[q * NB:(q + 1) * NB] = self.Vb[_block_rows(row0, lv, k), col0:col0 + NB]
        if first:                       # pivot block: unit lower from the stored strict lower part
            Vt[:NB] = np.tril(Vt[:NB], -1) + np.eye(NB)
        return Vt

    def _update_level(self, p, row0, li, lv, C, c0, c1, forward, ccol0=None):
        """Apply Q^T (forward) or Q (backward) of (panel p, level li) to C[:, c0:c1]."""
        col0 = p * NB
        for j in range(lv["nstrips"]):
            t0 = j * lv["s"]
            tiles = [t0 + i for i in range(lv["s"]) if t0 + i < lv["ntiles"]]
            piv = _block_rows(row0, lv, t0 * G)
            order = tiles if forward else tiles[::-1]
            Z = None if forward else C[piv, c0:c1].copy()
            for t in order:
                first = (t == t0)
                T = self.T[(p, li, t)]
                Top = T.T if forward else T
                Vt = self._tile_V(p, row0, col0, li, lv, t, first)
                blocks = self._tile_blocks(lv, t)
                Ct = np.zeros((G * NB, c1 - c0))
                for q, k in enumerate(blocks):
                    Ct[q * NB:(q + 1) * NB] = C[_block_rows(row0, lv, k), c0:c1]
                if first:
                    if not forward:
                        Ct[:NB] = Z
                    W = Top @ (Vt.T @ Ct)
                    Ct -= Vt @ W
                    Z = Ct[:NB].copy()
                else:
                    W = Top @ (Z + Vt.T @ Ct)
                    Z -= W
                    Ct -= Vt @ W
                for q, k in enumerate(blocks):
                    if first and q == 0:
                        continue
                    C[_block_rows(row0, lv, k), c0:c1] = Ct[q * NB:(q + 1) * NB]
            C[piv, c0:c1] = Z

    # ---- explicit Q1 (m x n) in place -------------------------------------------------------
    def form_q(self):
        Vb = self.Vb
        # zero the R entries right of each diagonal block (identity's off-diagonal part)
        for p in range(self.K):
            Vb[p * NB:(p + 1) * NB, (p + 1) * NB:] = 0.0
        for p in reversed(range(self.K)):
            row0 = col0 = p * NB
            levels = self.plans[p]
            Ptmp = np.zeros((self.Mrows, NB))
            Ptmp[row0:row0 + NB] = np.eye(NB)
            for li in reversed(range(len(levels))):
                lv = levels[li]
                self._update_level(p, row0, li, lv, Ptmp, 0, NB, forward=False)
                if col0 + NB < self.npad:
                    self._update_level(p, row0, li, lv, Vb, col0 + NB, self.npad, forward=False)
            Vb[row0:, col0:col0 + NB] = Ptmp[row0:]
            Vb[:row0, col0:col0 + NB] = 0.0
        return Vb[:self.m, :self.n]


def jacobi_svd_rows(R, tol=None, max_sweeps=40):
    """One-sided Jacobi on the ROWS of R (left rotations): R = Ur diag(S) Vt.  Device model."""
    n = R.shape[0]
    Gm = R.astype(float).copy()
    J = np.eye(n)
    tol = 2.0 * np.sqrt(n) * np.finfo(float).eps if tol is None else tol
    ne = n + (n & 1)
    for sweep in range(max_sweeps):
        rotated = 0
        for r in range(ne - 1):
            pairs = [(ne - 1, r)] + [((r + i) % (ne - 1), (r - i) % (ne - 1)) for i in range(1, ne // 2)]
            for (p, q) in pairs:
                if p >= n or q >= n:
                    continue
                if p > q:
                    p, q = q, p
                a = Gm[p] @ Gm[p]; b = Gm[q] @ Gm[q]; c = Gm[p] @ Gm[q]
                if abs(c) <= tol * np.sqrt(a * b) or c == 0.0:
                    continue
                rotated += 1
                zeta = (b - a) / (2.0 * c)
                t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                cs = 1.0 / np.sqrt(1.0 + t * t); sn = cs * t
                gp, gq = Gm[p].copy(), Gm[q].copy()
                Gm[p] = cs * gp - sn * gq; Gm[q] = sn * gp + cs * gq
                jp, jq = J[p].copy(), J[q].copy()
                J[p] = cs * jp - sn * jq; J[q] = sn * jp + cs * jq
        if rotated == 0:
            break
    s = np.linalg.norm(Gm, axis=1)
    order = np.argsort(-s, kind="stable")
    S = s[order]
    Vt = Gm[order] / np.where(S > 0, S, 1.0)[:, None]
    Ur = J[order].T
    return Ur, S, Vt, sweep + 1
