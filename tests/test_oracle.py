"""Pin the numpy oracle (oracle/pod_oracle.py) against (1) the golden vectors produced by the
reference's own Python sources (oracle/gen_golden.py) and (2) the reference's own C sources
compiled into oracle/_ref/libpylom_ref.so.  CPU only."""
import ctypes, glob, os
import numpy as np
import pytest

import pod_oracle as po
import synth

ALL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL_GOLDEN if not os.path.basename(p).startswith(("rsvd_", "dmd_"))]
DMD_GOLDEN = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("dmd_")]
RSVD_GOLDEN = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("rsvd_")]


def _split(A, P):
    return [A[slice(*po.worksplit(0, A.shape[0], r, P))] for r in range(P)]


def test_have_goldens():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_tsqr_svd_matches_reference_python(path):
    g = np.load(path)
    A = g["A"]
    for key in [k for k in g.files if k.startswith("tsqr_svd_P") and k.endswith("_S")]:
        P = int(key.split("_P")[1].split("_")[0])
        U, S, V = po.tsqr_svd(_split(A, P))
        U = np.vstack(U)
        m = po.compare_svd(g[f"tsqr_svd_P{P}_U"], g[key], g[f"tsqr_svd_P{P}_V"], U, S, V)
        assert m["sigma_rel"] <= 1e-13, (P, m)
        assert m["mode_min"] >= 1 - 1e-10 and m["vmode_min"] >= 1 - 1e-10, (P, m)
        # the oracle is the same LAPACK calls in the same order: expect bitwise S
        assert np.array_equal(S, g[key]), P
        assert np.abs(U.T @ U - np.eye(A.shape[1])).max() < 1e-13


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_pod_matches_reference_python(path):
    g = np.load(path)
    A = g["A"]
    for key in [k for k in g.files if k.startswith("pod_P") and k.endswith("_S")]:
        P = int(key.split("_P")[1].split("_")[0])
        shards = _split(A, P)
        U, S, V = po.pod_run(shards, remove_mean=True)
        assert np.allclose(S, g[key], rtol=0, atol=1e-13 * g[key][0])
        mean = np.concatenate([po.temporal_mean(X) for X in shards])
        assert np.array_equal(mean, g[f"pod_P{P}_mean"])
        Uc = np.vstack(U)
        Ur, Sr, Vr = po.truncate(Uc, S, V, r=1e-6)
        assert Sr.shape[0] == int(g[f"pod_P{P}_N"])
        Xr = po.reconstruct(Ur, Sr, Vr)
        assert np.abs(Xr - g[f"pod_P{P}_Xrec"]).max() <= 1e-12 * np.abs(A).max()
        Y = np.vstack([po.subtract_mean(X, po.temporal_mean(X)) for X in shards])
        assert abs(po.RMSE(Y, Xr) - float(g[f"pod_P{P}_rmse"])) <= 1e-12


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_truncation_rule(path):
    g = np.load(path)
    S = g["tsqr_svd_P1_S"]
    for r, N in zip(g["trunc_r"], g["trunc_N"]):
        assert po.compute_truncation_residual(S, float(r)) == int(N)


def test_worksplit_matches_the_reference_function():
    """tests/golden/aux/worksplit_ref.npz holds the outputs of the reference's own `worksplit` source
    (pyLOM/utils/parall.py:24-48, executed verbatim by oracle/gen_worksplit_golden.py): the oracle's restatement and the
    product's utils.worksplit must reproduce every entry (the golden row splits of gen_golden.py use the oracle's)."""
    tab = np.load(os.path.join(os.path.dirname(__file__), "golden", "aux", "worksplit_ref.npz"))["table"]
    assert len(tab) > 500
    import pyloworder_b200.utils as plu
    for i0, i1, rank, size, a, b in tab.tolist():
        assert tuple(po.worksplit(i0, i1, rank, size)) == (a, b), (i0, i1, rank, size)
        assert tuple(plu.worksplit(i0, i1, rank, size)) == (a, b), (i0, i1, rank, size)


AUX = os.path.join(os.path.dirname(__file__), "golden", "aux")


@pytest.mark.parametrize("name", ["ztsqr_svd_500x12", "stsqr_svd_800x20"])
def test_dtype_variants_match_reference_python(name):
    """complex128 / float32 tsqr_svd (ztsqr_svd / stsqr_svd, pyLOM/vmmath/src/svd.c:955-1010, 529-563): fixtures made by the
    reference's own dtype-generic Python tsqr_svd (oracle/gen_golden.py --dtypes).  The oracle is the same LAPACK calls in
    the same order, so it must reproduce them bitwise; the fixtures themselves are checked against LAPACK on the whole
    matrix (this is what the GPU tests of the real-embedding / widening paths compare with)."""
    g = np.load(os.path.join(AUX, name + ".npz"))
    A = g["A"]
    n = A.shape[1]
    eps = np.finfo(A.real.dtype).eps
    So = np.linalg.svd(A.astype(np.complex128 if np.iscomplexobj(A) else np.float64), compute_uv=False)
    for key in [k for k in g.files if k.endswith("_S")]:
        P = int(key[1:].split("_")[0])
        Ug, Sg, Vg = g[f"P{P}_U"], g[key], g[f"P{P}_V"]
        assert Ug.dtype == A.dtype and Sg.dtype == A.real.dtype
        assert np.abs(Sg - So).max() <= 20 * eps * So[0]
        assert np.abs((Ug * Sg) @ Vg - A).max() <= 50 * eps * np.abs(A).max() * np.sqrt(n)
        assert np.abs(Ug.conj().T @ Ug - np.eye(n)).max() <= 50 * eps
        U, S, V = po.tsqr_svd([np.ascontiguousarray(x) for x in _split(A, P)])
        assert np.array_equal(S, Sg) and np.array_equal(np.vstack(U), Ug) and np.array_equal(V, Vg), P


def test_worksplit_covers_range():
    for m, P in ((10, 3), (89351, 8), (7, 7), (5, 8), (1000, 1)):
        edges = [po.worksplit(0, m, r, P) for r in range(P)]
        assert edges[0][0] == 0
        for a, b in zip(edges[:-1], edges[1:]):
            assert a[1] == b[0] or m <= P
        assert max(e[1] for e in edges) == m


def test_synth_slices_reproducible():
    X = synth.snapshots(1000, 16, 2021)
    assert np.array_equal(X[300:450], synth.snapshots(1000, 16, 2021, 300, 450))
    assert np.linalg.matrix_rank(X) == 16


# ---- the reference's own C sources -------------------------------------------------------------
REFLIB = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libpylom_ref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built (make -C oracle ref)")
dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(dp)


@needs_ref
def test_oracle_vs_reference_c_tsqr_svd():
    lib = ctypes.CDLL(REFLIB)
    for m, n, seed in ((500, 12, 3), (2000, 64, 4), (1800, 151, 2021)):
        A = synth.snapshots(m, n, seed) if n == 151 else synth.random_matrix(m, n, seed, cond=1e6)
        U = np.zeros((m, n)); S = np.zeros(n); V = np.zeros((n, n))
        info = lib.dtsqr_svd(_p(U), _p(S), _p(V), _p(A.copy()), ctypes.c_int(m), ctypes.c_int(n))
        assert info == 0
        Uo, So, Vo = po.tsqr_svd(A)
        mtr = po.compare_svd(U, S, V, Uo, So, Vo)
        assert mtr["sigma_rel"] <= 1e-13 and mtr["sigma_rel_each"] <= 1e-9, mtr
        assert mtr["mode_min"] >= 1 - 1e-9, mtr


@needs_ref
def test_oracle_vs_reference_c_averaging():
    lib = ctypes.CDLL(REFLIB)
    X = synth.snapshots(777, 37, 5)
    mean = np.zeros(777); Y = np.zeros_like(X)
    lib.dtemporal_mean(_p(mean), _p(X), ctypes.c_int(777), ctypes.c_int(37))
    lib.dsubtract_mean(_p(Y), _p(X), _p(mean), ctypes.c_int(777), ctypes.c_int(37))
    assert np.abs(mean - po.temporal_mean(X)).max() <= 4e-16 * np.abs(X).max() * 37
    assert np.abs(Y - po.subtract_mean(X, mean)).max() == 0.0


@pytest.mark.parametrize("path", RSVD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_randomized_svd_against_reference_golden(path):
    """randomized_qr / randomized_svd / POD.run(randomized=True) of the oracle against the reference's own Python
    (one rank), and the P = 2, 3 simulated-rank runs against the same fixture (the sketch is the same on all ranks)."""
    g = np.load(path)
    A, r, q, sk = g["A"], int(g["r"]), int(g["q"]), int(g["seed"])
    assert np.array_equal(po.sketch_matrix(A.shape[1], r, sk), g["omega"])
    assert len(RSVD_GOLDEN) >= 3
    for P in (1, 2, 3):
        blocks = [A[slice(*po.worksplit(0, A.shape[0], k, P))] for k in range(P)]
        Q, B = po.randomized_qr(blocks, r, q, sk)
        Q = np.vstack(Q)
        assert np.abs(Q.T @ Q - np.eye(r)).max() <= 1e-13
        assert np.abs(Q.T @ A - B).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
        assert np.abs(g["Q"] @ (g["Q"].T @ Q) - Q).max() <= 1e-9          # same range as the reference's Q
        U, S, V = po.randomized_svd(blocks, r, q, sk)
        mt = po.compare_svd(g["U"], g["S"], g["V"], np.vstack(U), S, V)
        assert mt["sigma_rel"] <= 1e-12 and mt["mode_min"] >= 1 - 1e-8 and mt["vmode_min"] >= 1 - 1e-8, mt
    U, S, V = po.pod_run(A, remove_mean=True, randomized=True, r=r, q=q, seed=sk)
    mt = po.compare_svd(g["pod_U"], g["pod_S"], g["pod_V"], U, S, V)
    assert mt["sigma_rel"] <= 1e-12 and mt["mode_min"] >= 1 - 1e-8, mt


@pytest.mark.parametrize("path", RSVD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_streaming_randomized_qr_against_reference_golden(path):
    """init_qr_streaming / update_qr_streaming: one rank reproduces the reference bit for bit (same LAPACK, same
    MT19937 stream).  The update ADDS two sketches whose column signs come from the QR, so its result depends on the
    QR's sign convention (the reference's own P = 2 run differs from its P = 1 run); for P = 2 only the identities
    the algorithm guarantees are checked."""
    g = np.load(path)
    A, r, q, sk, n1 = g["A"], int(g["r"]), int(g["q"]), int(g["seed"]), int(g["st_n1"])
    A1, A2 = np.ascontiguousarray(A[:, :n1]), np.ascontiguousarray(A[:, n1:])
    Q1, B1, Y1 = po.init_qr_streaming(A1, r, q, sk)
    assert np.array_equal(Q1, g["st_Q1"]) and np.array_equal(B1, g["st_B1"]) and np.array_equal(Y1, g["st_Y1"])
    Q2, B2, Y2 = po.update_qr_streaming(A2, Q1, B1, Y1, r, q)
    assert np.array_equal(Q2, g["st_Q2"]) and np.array_equal(B2, g["st_B2"]) and np.array_equal(Y2, g["st_Y2"])
    P = 2
    cut = lambda X: [X[slice(*po.worksplit(0, X.shape[0], k, P))] for k in range(P)]
    Q1, B1, Y1 = po.init_qr_streaming(cut(A1), r, q, sk)
    Q2, B2, Y2 = po.update_qr_streaming(cut(A2), Q1, B1, Y1, r, q)
    Q1, Q2, Y2 = np.vstack(Q1), np.vstack(Q2), np.vstack(Y2)
    assert B2.shape == (r, A.shape[1])
    assert np.abs(Q2.T @ Q2 - np.eye(r)).max() <= 1e-13
    assert np.abs(Q2 @ (Q2.T @ Y2) - Y2).max() <= 1e-10 * np.abs(Y2).max()
    assert np.abs(B2[:, n1:] - Q2.T @ A2).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
    assert np.abs(B2[:, :n1] - (Q2.T @ Q1) @ B1).max() <= 1e-12 * np.abs(A).max() * A.shape[1]


def dmd_invariants(g, muR, muI, Phi, b, key="P1", mu_tol=1e-10, amp_tol=1e-6):
    """DMD results are unique up to a unit complex factor c_k per mode (Phi_k -> c_k Phi_k, b_k -> b_k / c_k: LAPACK
    only fixes the norm of an eigenvector) -- compare eigenvalues, |b|, the products Phi_k b_k and the reconstruction.
    The amplitudes solve a Vandermonde-Gram system with cond ~ 1e8, hence the looser tolerance."""
    assert np.abs(muR - g[key + "_muReal"]).max() <= mu_tol and np.abs(muI - g[key + "_muImag"]).max() <= mu_tol
    bref, Pref = g[key + "_b"], g[key + "_Phi"]
    assert np.abs(np.abs(b) - np.abs(bref)).max() <= amp_tol * np.abs(bref).max()
    assert np.abs(Phi * b - Pref * bref).max() <= amp_tol * np.abs(Pref * bref).max()


@pytest.mark.parametrize("path", DMD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_dmd_against_reference_golden(path):
    g = np.load(path)
    X, r, dt = g["X"], float(g["r"]), float(g["dt"])
    assert len(DMD_GOLDEN) >= 2
    muR, muI, Phi, b = po.dmd_run(X, r)
    assert np.array_equal(muR, g["P1_muReal"]) and np.array_equal(Phi, g["P1_Phi"])      # same LAPACK, same statements
    dmd_invariants(g, muR, muI, Phi, b)
    assert np.all(muI[0::2] >= 0)                       # positive member of every conjugate pair first
    delta, omega = po.dmd_frequency_damping(muR, muI, dt)
    assert np.allclose(delta, g["delta"], rtol=0, atol=1e-12) and np.allclose(omega, g["omega"], rtol=0, atol=1e-12)
    t = np.arange(X.shape[1], dtype=np.double)
    Xd = po.dmd_reconstruction_jovanovic(Phi, muR, muI, t, b)
    assert np.abs(Xd - g["X_DMD"]).max() <= 1e-6 * np.abs(g["X_DMD"]).max()
    for key in [k[:-7] for k in g.files if k.endswith("_muReal") and k != "P1_muReal"]:
        P = int(key[1:])
        bl = [X[slice(*po.worksplit(0, X.shape[0], k, P))] for k in range(P)]
        muR, muI, Phi, b = po.dmd_run(bl, r)
        dmd_invariants(g, muR, muI, np.vstack(Phi), b, key)          # the reference's own P-rank run
        dmd_invariants(g, muR, muI, np.vstack(Phi), b, "P1")


def test_ref_ranks_processes_match_simulated_butterfly():
    """oracle/ref_ranks.py runs the reference's P-rank tsqr_svd as P real processes (the bench's CPU arm);
    its numpy-kernel variant must reproduce the level-synchronous simulation bit for bit in S, and its
    reference-C-kernel variant (oracle/_ref) to rounding."""
    import ref_ranks
    m, n = 3000, 24
    X = synth.snapshots(m, n, 2022)
    for P in (2, 3):
        shards = [X[slice(*po.worksplit(0, m, k, P))] for k in range(P)]
        Uo, So, Vo = po.tsqr_svd(shards)
        r = ref_ranks.run(P, m, n, 2022, steps=1, warmup=0, keep=True, use_ref=False)
        assert all(np.array_equal(r["S"][0], s) for s in r["S"])          # identical on all ranks
        assert np.array_equal(r["S"][0], So)
        mt = po.compare_svd(np.vstack(Uo), So, Vo, np.vstack(r["U"]), r["S"][0], r["VT"][0])
        assert mt["mode_min"] >= 1 - 1e-12
    if os.path.exists(ref_ranks.REF_SO):
        r = ref_ranks.run(2, m, n, 2022, steps=1, warmup=0, keep=True, use_ref=True)
        assert r["kind"] == "reference"
        shards = [X[slice(*po.worksplit(0, m, k, 2))] for k in range(2)]
        Uo, So, Vo = po.tsqr_svd(shards)
        mt = po.compare_svd(np.vstack(Uo), So, Vo, np.vstack(r["U"]), r["S"][0], r["VT"][0])
        assert mt["sigma_rel"] <= 1e-13 and mt["mode_min"] >= 1 - 1e-10
