"""The numpy model of the device algorithm (tests/model_caqr.py) against LAPACK -- CPU only."""
import numpy as np
import pytest

import model_caqr as mc
import model_tsqr_small as ms
import synth


@pytest.mark.parametrize("m,n,kind", [(40, 40, "rand"), (700, 24, "synth"), (1500, 64, "rand"),
                                      (5000, 70, "cond"), (9000, 33, "center"), (20000, 8, "rand")])
def test_caqr_model(m, n, kind, monkeypatch):
    if kind == "synth":
        A = synth.snapshots(m, n, 1)
    elif kind == "cond":
        A = synth.random_matrix(m, n, 2, cond=1e10)
    elif kind == "center":
        A = synth.snapshots(m, n, 3); A = A - A.mean(axis=1, keepdims=True)
    else:
        A = synth.random_matrix(m, n, 4)
    # force deep trees on small inputs
    monkeypatch.setattr(mc, "plan_levels", lambda m_act: _plan(m_act))
    f = mc.CAQR(A)
    R = f.factor()
    Rref = np.linalg.qr(A, mode="r")
    sc = np.abs(A).max()
    assert np.abs(np.abs(R) - np.abs(Rref)).max() <= 1e-12 * sc * n
    Q = f.form_q().copy()
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14
    assert np.abs(Q @ R - A).max() <= 1e-13 * sc * n


def _plan(m_act):
    levels = []
    nblk, bs = -(-m_act // mc.NB), mc.NB
    while True:
        ntiles = -(-nblk // mc.G)
        s = 3 if ntiles > 6 else min(ntiles, 2)
        nstrips = -(-ntiles // s)
        levels.append(dict(nblk=nblk, bs=bs, ntiles=ntiles, s=s, nstrips=nstrips))
        if nstrips == 1:
            return levels
        nblk, bs = nstrips, bs * mc.G * s


@pytest.mark.parametrize("n,cond", [(5, 1e3), (32, 1e8), (77, 1e14)])
def test_jacobi_model(n, cond):
    A = synth.random_matrix(4 * n, n, 9, cond=cond)
    R = np.linalg.qr(A, mode="r")
    Ur, S, Vt, sweeps = mc.jacobi_svd_rows(R)
    Sref = np.linalg.svd(R, compute_uv=False)
    assert np.abs(S - Sref).max() <= 1e-14 * Sref[0]
    assert np.max(np.abs(S - Sref) / Sref) <= 1e-9
    assert np.abs(Ur.T @ Ur - np.eye(n)).max() < 1e-13
    assert np.abs(Vt @ Vt.T - np.eye(n)).max() < 1e-12
    assert np.abs((Ur * S) @ Vt - R).max() <= 1e-13 * Sref[0]
    assert sweeps < 20


@pytest.mark.parametrize("m,n,NP,kind", [(9000, 64, 64, "rand"), (12000, 50, 64, "center"), (7000, 32, 32, "rand"),
                                         (9000, 20, 32, "cond"), (5300, 64, 64, "rand"), (20000, 64, 64, "def")])
def test_tsqr_small_model(m, n, NP, kind):
    """The two-pass tile TSQR of csrc/tsqr_small.cu (dense head + left-looking structured tiles with an accumulated
    T, zero-block back-multiply) against LAPACK: several strips, a partial last tile, rank-deficient input."""
    rng = np.random.default_rng(m + n)
    if kind == "center":
        A = synth.snapshots(m, n, 3); A = A - A.mean(axis=1, keepdims=True)
    elif kind == "cond":
        A = synth.random_matrix(m, n, 2, cond=1e12)
    else:
        A = rng.standard_normal((m, n))
        if kind == "def":
            A[:, 5] = 0.0; A[:, 17] = A[:, 3]
    Q, R = ms.qr(A, NP, nstrips_target=5, min_tiles=2)
    sc = np.abs(A).max()
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14
    assert np.abs(Q @ R - A).max() <= 1e-13 * sc * n
    # strips: NP head rows + whole 512-row tiles, only the last strip may end with a partial tile
    strips = ms.plan(m, NP, 5, 2)
    assert sum(NP + (nt - 1) * ms.TB + last for _, nt, last in strips) == m
    assert all(last == ms.TB for _, _, last in strips[:-1])
