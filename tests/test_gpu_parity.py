"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the committed golden vectors produced by
the reference's own sources.  Tolerances are BASELINE.json's: singular values 1e-10 relative, modes
|<u_ref,u>| >= 1 - 1e-8 for well separated modes, reconstruction RMSE equal to 1e-10."""
import ctypes, glob, os
import numpy as np
import pytest
import torch

import pod_oracle as po
import synth

pytestmark = pytest.mark.gpu
ALL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL_GOLDEN if not os.path.basename(p).startswith(("rsvd_", "dmd_"))]
DMD_GOLDEN = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("dmd_")]
RSVD_GOLDEN = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("rsvd_")]

SIG_TOL = 1e-10
MODE_TOL = 1 - 1e-8
RMSE_TOL = 1e-10


@pytest.fixture(scope="module")
def pl():
    import pyloworder_b200
    assert torch.cuda.is_available()
    return pyloworder_b200


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def assert_svd_parity(ref, got, n_expected=None):
    mt = po.compare_svd(*ref, *got)
    assert mt["sigma_rel"] <= SIG_TOL, mt
    assert mt["sigma_rel_each"] <= SIG_TOL, mt
    assert mt["mode_min"] >= MODE_TOL, mt
    assert mt["vmode_min"] >= MODE_TOL, mt
    return mt


# ---- golden vectors from the reference's own Python sources --------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_tsqr_svd_against_reference_golden(pl, path):
    g = np.load(path)
    A = g["A"]
    U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(A))]
    assert U.shape == A.shape and S.shape == (A.shape[1],) and V.shape == (A.shape[1],) * 2
    assert_svd_parity((g["tsqr_svd_P1_U"], g["tsqr_svd_P1_S"], g["tsqr_svd_P1_V"]), (U, S, V))
    n = A.shape[1]
    assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-12
    assert np.abs((U * S) @ V - A).max() <= 1e-12 * max(1.0, np.abs(A).max())


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_pod_pipeline_against_reference_golden(pl, path):
    g = np.load(path)
    A = g["A"]
    Ad = dev(A)
    U, S, V = pl.POD.run(Ad, remove_mean=True)
    assert torch.equal(Ad, dev(A)), "POD.run must not modify X"
    assert_svd_parity((g["pod_P1_U"], g["pod_P1_S"], g["pod_P1_V"]), (host(U), host(S), host(V)))
    Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=1e-6)
    assert Sr.shape[0] == int(g["pod_P1_N"])
    Xr = pl.POD.reconstruct(Ur, Sr, Vr)
    assert np.abs(host(Xr) - g["pod_P1_Xrec"]).max() <= 1e-10 * max(1.0, np.abs(A).max())
    mean = pl.math.temporal_mean(Ad)
    assert np.abs(host(mean) - g["pod_P1_mean"]).max() <= 4e-16 * A.shape[1] * np.abs(A).max()
    Y = pl.math.subtract_mean(Ad, mean)
    rm = pl.math.RMSE(Y, Xr)
    assert abs(rm - float(g["pod_P1_rmse"])) <= RMSE_TOL


# ---- seeded inputs against the oracle -------------------------------------------------------------
CASES = [
    (64, 64, "rand"),          # m == n
    (33, 1, "rand"),           # single column
    (1000, 2, "rand"),
    (4097, 31, "rand"),        # odd n, ragged m
    (2500, 32, "synth"),
    (40000, 64, "synth"),      # cfg5 twin (n = 64)
    (89351, 151, "synth"),     # cfg1 at full size
    (60000, 256, "synth"),     # cfg3 twin (n = 256)
    (20000, 512, "synth"),     # cfg2 twin (n = 512)
    (12000, 999, "synth"),     # cfg4 twin (n = 999)
    (30000, 100, "cond1e12"),
]


@pytest.mark.parametrize("m,n,kind", CASES, ids=lambda v: str(v))
def test_tsqr_svd_against_oracle(pl, m, n, kind):
    if kind == "rand":
        A = synth.random_matrix(m, n, 17)
    elif kind == "cond1e12":
        A = synth.random_matrix(m, n, 5, cond=1e12)
    else:
        A = synth.snapshots(m, n, 2021)
    U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(A))]
    Uo, So, Vo = po.tsqr_svd(A)
    assert_svd_parity((Uo, So, Vo), (U, S, V))
    assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-12
    assert np.linalg.norm((U * S) @ V - A) / np.linalg.norm(A) <= 1e-13 * np.sqrt(n) + 1e-14
    assert np.all(np.diff(S) <= 0)


@pytest.mark.parametrize("m,n", [(89351, 151), (50000, 256), (7000, 20)])
def test_pod_run_remove_mean_against_oracle(pl, m, n):
    X = synth.snapshots(m, n, 2023, nvars=1)
    U, S, V = pl.POD.run(dev(X), remove_mean=True)
    Uo, So, Vo = po.pod_run(X, remove_mean=True)
    assert_svd_parity((Uo, So, Vo), (host(U), host(S), host(V)))
    # centred data is rank deficient (rows sum to zero): the basis must still be orthonormal
    Uh = host(U)
    assert np.abs(Uh.T @ Uh - np.eye(n)).max() <= 1e-12
    # ... and so must V (LAPACK completes the null direction of V^T to an orthonormal set; so does svd_complete_kernel)
    Vh = host(V)
    assert np.abs(Vh @ Vh.T - np.eye(n)).max() <= 1e-12
    for r in (1e-6, 5, -0.9):
        Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=r)
        Xr = host(pl.POD.reconstruct(Ur, Sr, Vr))
        Xo = po.reconstruct(*po.truncate(Uo, So, Vo, r=r))
        assert Ur.shape[1] == Xo.shape[1] * 0 + po.truncate(Uo, So, Vo, r=r)[1].shape[0]
        Y = po.subtract_mean(X, po.temporal_mean(X))
        assert abs(po.RMSE(Y, Xr) - po.RMSE(Y, Xo)) <= RMSE_TOL
    U2, S2, V2 = pl.POD.run(dev(X), remove_mean=False)
    Uo2, So2, Vo2 = po.pod_run(X, remove_mean=False)
    assert_svd_parity((Uo2, So2, Vo2), (host(U2), host(S2), host(V2)))


def test_averaging_kernels(pl):
    for m, n in ((1, 1), (5, 3), (1000, 37), (513, 64), (4000, 512), (100, 1000), (70000, 151)):
        X = synth.snapshots(max(m, 2), n, 3)[:m]
        mean = pl.math.temporal_mean(dev(X))
        ref = po.temporal_mean(X)
        assert np.abs(host(mean) - ref).max() <= 4e-16 * n * np.abs(X).max()
        Y = pl.math.subtract_mean(dev(X), dev(ref))
        assert np.array_equal(host(Y), po.subtract_mean(X, ref))            # one subtraction: bit exact
        Yn = pl.math.subtract_mean(X, ref)                                   # numpy in -> numpy out
        assert isinstance(Yn, np.ndarray) and np.array_equal(Yn, host(Y))
        # the documented way to add the mean back (docs notebook: subtract_mean(X_POD, -1*mean))
        Z = pl.math.subtract_mean(Y, -1 * dev(ref))
        assert np.abs(host(Z) - X).max() <= 2e-16 * np.abs(X).max() * 2


def test_variance_normalisation(pl):
    X = synth.snapshots(5000, 33, 8)
    mean = po.temporal_mean(X)
    var = pl.math.temporal_variance(dev(X), dev(mean))
    assert var.shape == (5000, 1)
    assert np.abs(host(var) - po.temporal_variance(X, mean)).max() <= 1e-14 * np.abs(po.temporal_variance(X, mean)).max()
    Y = pl.math.norm_variance(dev(X), dev(mean), dev(po.temporal_variance(X, mean)))
    assert np.array_equal(host(Y), po.norm_variance(X, mean, po.temporal_variance(X, mean)))
    U, S, V = pl.POD.run(dev(X), remove_mean=True, divide_variance=True)
    Uo, So, Vo = po.pod_run(X, remove_mean=True, divide_variance=True)
    assert_svd_parity((Uo, So, Vo), (host(U), host(S), host(V)))
    # fused C entry point (centering + variance normalisation inside the factorisation copy)
    from pyloworder_b200 import _lib, _dev
    L = _lib.lib()
    m, n = X.shape
    Xd = dev(X)
    R = torch.empty((n, n), dtype=torch.float64, device="cuda"); mu = torch.empty(m, dtype=torch.float64, device="cuda"); vv = torch.empty_like(mu)
    _, wp, wb = _dev.workspace(L.pl_qr_workspace_bytes(m, n), "t", Xd.device)
    assert L.pl_qr_factor_var_f64(R.data_ptr(), mu.data_ptr(), vv.data_ptr(), Xd.data_ptr(), m, n, wp, wb, None) == 0
    Rr = np.linalg.qr(po.norm_variance(X, mean, po.temporal_variance(X, mean)), mode="r")
    assert np.abs(np.abs(host(R)) - np.abs(Rr)).max() <= 1e-11 * np.abs(Rr).max()
    assert np.abs(host(vv) - po.temporal_variance(X, mean)[:, 0]).max() <= 1e-14 * host(vv).max()


def test_matmul_vecmat_reconstruct(pl):
    rng = np.random.default_rng(0)
    for m, n, k in ((1000, 64, 64), (777, 151, 151), (5000, 512, 512), (300, 40, 7), (129, 33, 100), (1, 1, 1), (257, 999, 999)):
        A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
        C = host(pl.math.matmul(dev(A), dev(B)))
        assert np.abs(C - A @ B).max() <= 1e-13 * k * 10
    v = rng.standard_normal(37); A = rng.standard_normal((37, 50))
    assert np.array_equal(host(pl.math.vecmat(dev(v), dev(A))), po.vecmat(v, A))
    # reconstruct on truncated (strided) views
    U = rng.standard_normal((900, 40)); S = np.abs(rng.standard_normal(40)); V = rng.standard_normal((40, 60))
    Ud, Sd, Vd = dev(U), dev(S), dev(V)
    X = host(pl.POD.reconstruct(Ud[:, :7], Sd[:7], Vd[:7, :]))
    assert np.abs(X - po.reconstruct(U[:, :7], S[:7], V[:7, :])).max() <= 1e-12


def test_energy_and_extract_modes(pl):
    X = synth.snapshots(9000, 20, 4, nvars=3)
    U, S, V = pl.POD.run(dev(X), remove_mean=False)
    Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=4)
    Xr = pl.POD.reconstruct(Ur, Sr, Vr)
    e = pl.math.energy(dev(X), Xr)
    assert abs(e - po.energy(X, host(Xr))) <= 1e-13
    assert abs((1 - e) - po.RMSE(X, host(Xr)) ** 2) <= 1e-13
    Uh = host(U)
    for ivar in (1, 2, 3):
        for modes, reshape in (([], True), ([1, 3, 4], False), ([2], True)):
            got = pl.POD.extract_modes(U, ivar, 3000, modes=modes, reshape=reshape)
            ref = po.extract_modes(Uh, ivar, 3000, modes=modes, reshape=reshape)
            assert np.array_equal(host(got), ref)
            assert np.array_equal(pl.POD.extract_modes(Uh, ivar, 3000, modes=modes, reshape=reshape), ref)


def test_qr_and_svd_entry_points(pl):
    A = synth.random_matrix(3000, 45, 1, cond=1e6)
    Q, R = [host(t) for t in pl.math.qr(dev(A))]
    assert np.allclose(np.tril(R, -1), 0)
    assert np.abs(Q.T @ Q - np.eye(45)).max() <= 1e-13
    assert np.abs(Q @ R - A).max() <= 1e-13 * np.abs(A).max() * 45
    Rr = np.linalg.qr(A, mode="r")
    assert np.abs(np.abs(R) - np.abs(Rr)).max() <= 1e-12 * np.abs(Rr).max()
    Qt, Rt = [host(t) for t in pl.math.tsqr(dev(A))]
    assert np.abs(Qt @ Rt - A).max() <= 1e-13 * np.abs(A).max() * 45
    U, S, V = [host(t) for t in pl.math.svd(dev(R))]
    So = np.linalg.svd(R, compute_uv=False)
    assert np.max(np.abs(S - So) / So) <= 1e-10


@pytest.mark.parametrize("container", ["torch", "numpy"])
def test_float32_inputs(pl, container):
    """float32 callers (the reference's `real` fused type covers float: stsqr_svd, pyLOM/vmmath/src/svd.c:529-563): inputs are
    widened on the device, the fp64 path runs, results come back as float32.  Compared with the oracle run on the
    same (float32-valued) data; tolerances are single-precision ones."""
    X32 = synth.snapshots(20000, 48, 77).astype(np.float32)
    Xin = torch.from_numpy(X32).cuda() if container == "torch" else X32
    U, S, V = pl.POD.run(Xin, remove_mean=True)
    assert (U.dtype, S.dtype, V.dtype) == ((torch.float32,) * 3 if container == "torch" else (np.float32,) * 3)
    Uo, So, Vo = po.pod_run(X32.astype(np.float64), remove_mean=True)
    Uh, Sh, Vh = [np.asarray(host(t), dtype=np.float64) for t in (U, S, V)]
    assert np.abs(Sh - So).max() <= 2e-7 * So[0]
    big = So / So[0] >= 1e-3
    gaps = np.minimum(np.r_[np.inf, -np.diff(So)], np.r_[-np.diff(So), np.inf]) / So[0] >= 1e-3
    sel = big & gaps
    assert np.abs(np.einsum("ik,ik->k", Uo, Uh))[sel].min() >= 1 - 1e-5
    Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=5)
    Xr = pl.POD.reconstruct(Ur, Sr, Vr)
    assert Xr.dtype == U.dtype and tuple(Xr.shape) == X32.shape
    mean = pl.math.temporal_mean(Xin)
    assert mean.dtype == U.dtype and np.abs(np.asarray(host(mean), dtype=np.float64) - X32.astype(np.float64).mean(1)).max() <= 1e-6
    with pytest.raises(NotImplementedError):
        pl.math.tsqr_svd(torch.zeros((64, 4), dtype=torch.complex64, device="cuda"))


@pytest.mark.parametrize("case", ["generic", "repeated", "deficient", "numpy"])
def test_complex_tsqr_svd(pl, case):
    """complex128 tsqr_svd (the call of SPOD, pyLOM/SPOD/wrapper.py:83; ztsqr_svd src/svd.c:955-1010) through the real
    embedding: singular values against LAPACK, modes up to a complex phase, U^H U = I, A = U S V^H."""
    rng = np.random.default_rng(5)
    m, n = 6000, 24
    if case in ("generic", "numpy"):
        A = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
        A *= 10.0 ** (-3.0 * np.arange(n) / n)
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))
        W, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        sv = np.linspace(3.0, 1.0, n)
        sv[5] = sv[4]; sv[11] = sv[10] = sv[9]                       # coinciding complex singular values
        if case == "deficient":
            sv[-3:] = 0.0
        A = (Q * sv) @ W.conj().T
    Ain = A if case == "numpy" else torch.from_numpy(A).cuda()
    U, S, VH = pl.math.tsqr_svd(Ain)
    U, S, VH = [np.asarray(host(t)) for t in (U, S, VH)]
    assert U.shape == (m, n) and S.shape == (n,) and VH.shape == (n, n) and np.iscomplexobj(U) and np.iscomplexobj(VH)
    So = np.linalg.svd(A, compute_uv=False)
    assert np.abs(S - So).max() <= 1e-12 * So[0]
    assert np.abs(VH @ VH.conj().T - np.eye(n)).max() <= 1e-11
    assert np.abs((U * S) @ VH - A).max() <= 1e-11 * np.abs(A).max()
    nz = So > 1e-9 * So[0]
    G = U[:, nz].conj().T @ U[:, nz]
    assert np.abs(G - np.eye(int(nz.sum()))).max() <= 1e-10
    if case in ("generic", "numpy"):
        Uo = np.linalg.svd(A, full_matrices=False)[0]
        assert np.abs(np.einsum("ik,ik->k", Uo.conj(), U)).min() >= 1 - 1e-8


def test_complex_and_float32_against_reference_fixtures(pl):
    """tests/golden/aux/{ztsqr_svd_500x12,stsqr_svd_800x20}.npz: outputs of the reference's own (dtype-generic) Python
    tsqr_svd on complex128 / float32 input (oracle/gen_golden.py --dtypes; the C twins are ztsqr_svd / stsqr_svd,
    pyLOM/vmmath/src/svd.c:955-1010, 529-563).  complex128: singular values to 1e-12, modes up to a unit complex factor.
    float32: results come back as float32; tolerances are single-precision ones (the reference computed in float32)."""
    aux = os.path.join(os.path.dirname(__file__), "golden", "aux")
    g = np.load(os.path.join(aux, "ztsqr_svd_500x12.npz"))
    A, Ug, Sg, Vg = g["A"], g["P1_U"], g["P1_S"], g["P1_V"]
    n = A.shape[1]
    U, S, VH = [np.asarray(host(t)) for t in pl.math.tsqr_svd(torch.from_numpy(A).cuda())]
    assert U.dtype == np.complex128 and S.dtype == np.float64 and VH.dtype == np.complex128
    assert np.abs(S - Sg).max() <= 1e-12 * Sg[0]
    assert np.abs(np.einsum("ik,ik->k", Ug.conj(), U)).min() >= 1 - 1e-8
    assert np.abs(np.einsum("kj,kj->k", Vg.conj(), VH)).min() >= 1 - 1e-8
    assert np.abs((U * S) @ VH - A).max() <= 1e-11 * np.abs(A).max()
    assert np.abs(U.conj().T @ U - np.eye(n)).max() <= 1e-10
    g = np.load(os.path.join(aux, "stsqr_svd_800x20.npz"))
    A, Ug, Sg, Vg = g["A"], g["P1_U"].astype(np.float64), g["P1_S"].astype(np.float64), g["P1_V"].astype(np.float64)
    n = A.shape[1]
    out = pl.math.tsqr_svd(torch.from_numpy(A).cuda())
    assert all(t.dtype == torch.float32 for t in out)
    U, S, V = [np.asarray(host(t), dtype=np.float64) for t in out]
    assert np.abs(S - Sg).max() <= 1e-6 * Sg[0]
    assert np.abs(np.einsum("ik,ik->k", Ug, U)).min() >= 1 - 1e-5
    assert np.abs(np.einsum("kj,kj->k", Vg, V)).min() >= 1 - 1e-5
    assert np.abs((U * S) @ V - A).max() <= 1e-5 * np.abs(A).max()
    assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-5


def test_exactly_rank_deficient_input(pl):
    """All-zero trailing snapshot columns give exactly zero singular values; U and V must stay orthonormal
    (LAPACK returns an arbitrary orthonormal completion, so only the invariants are compared)."""
    A = synth.random_matrix(4000, 40, 3)
    A[:, 37:] = 0.0
    A[:, 5] = A[:, 4]                      # and one exactly duplicated column
    U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(A))]
    So = np.linalg.svd(A, compute_uv=False)
    assert np.abs(S - So).max() <= 1e-13 * So[0]
    assert np.abs(U.T @ U - np.eye(40)).max() <= 1e-12
    assert np.abs(V @ V.T - np.eye(40)).max() <= 1e-12
    assert np.abs((U * S) @ V - A).max() <= 1e-12 * np.abs(A).max()
    Z = np.zeros((300, 7))
    U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(Z))]
    assert np.all(S == 0) and np.abs(V @ V.T - np.eye(7)).max() <= 1e-14 and np.abs(U.T @ U - np.eye(7)).max() <= 1e-14


def test_error_behaviour(pl):
    with pytest.raises(ValueError, match="at least n rows"):
        pl.math.tsqr_svd(dev(np.zeros((3, 5))))
    with pytest.raises(NotImplementedError):
        pl.math.tsqr_svd(torch.zeros((10, 2), dtype=torch.complex64, device="cuda"))     # float32 is accepted (widened)
    with pytest.raises(ValueError, match="1 <= r <= n"):
        pl.POD.run(dev(np.ones((10, 2))), randomized=True, r=3)
    from pyloworder_b200 import _lib
    L = _lib.lib()
    rc = L.pl_tsqr_svd_f64(None, None, None, None, 3, 5, None, 0, None)
    assert rc < 0


def test_c_abi_host_entry_point(pl):
    """pl_tsqr_svd_host_f64: same arguments as the reference's dtsqr_svd, host pointers."""
    from pyloworder_b200 import _lib
    L = _lib.lib()
    m, n = 5000, 48
    A = synth.snapshots(m, n, 9)
    U = np.zeros((m, n)); S = np.zeros(n); V = np.zeros((n, n))
    rc = L.pl_tsqr_svd_host_f64(U.ctypes.data, S.ctypes.data, V.ctypes.data, A.ctypes.data, m, n)
    assert rc == 0, L.pl_last_error()
    assert_svd_parity(po.tsqr_svd(A), (U, S, V))
    reflib = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libpylom_ref.so")
    if os.path.exists(reflib):   # and against the reference's own compiled C, same call shape
        R = ctypes.CDLL(reflib)
        dp = ctypes.POINTER(ctypes.c_double)
        U2 = np.zeros((m, n)); S2 = np.zeros(n); V2 = np.zeros((n, n))
        assert R.dtsqr_svd(U2.ctypes.data_as(dp), S2.ctypes.data_as(dp), V2.ctypes.data_as(dp), A.ctypes.data_as(dp),
                           ctypes.c_int(m), ctypes.c_int(n)) == 0
        assert_svd_parity((U2, S2, V2), (U, S, V))


@pytest.mark.parametrize("m,n,chunks", [(20000, 96, 3), (20000, 100, 5), (9001, 33, 7), (4000, 64, 64)])
def test_c_abi_host_pipeline_chunked(pl, monkeypatch, m, n, chunks):
    """The host entry point as a multi-chunk (two-level TSQR) pipeline: padded and unpadded widths, ragged last
    chunk, more chunks requested than the 4n-rows rule allows."""
    from pyloworder_b200 import _lib
    L = _lib.lib()
    monkeypatch.setenv("PL_HOST_CHUNKS", str(chunks))
    A = synth.snapshots(m, n, 21)
    U = np.zeros((m, n)); S = np.zeros(n); V = np.zeros((n, n))
    for _ in range(2):      # second call reuses the cached device buffers
        rc = L.pl_tsqr_svd_host_f64(U.ctypes.data, S.ctypes.data, V.ctypes.data, A.ctypes.data, m, n)
        assert rc == 0, L.pl_last_error()
        assert_svd_parity(po.tsqr_svd(A), (U, S, V))
        assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-12
    L.pl_host_cache_free()


@pytest.mark.parametrize("chunks", [1, 4])
def test_c_abi_host_phase_calls(pl, monkeypatch, chunks):
    """The multi-rank host recipe (factor -> exchange of R -> stack SVD -> apply) run for one rank with HOST pointers
    for every argument, and the misuse cases (apply without factor, mismatched sizes)."""
    from pyloworder_b200 import _lib
    L = _lib.lib()
    monkeypatch.setenv("PL_HOST_CHUNKS", str(chunks))
    m, n = 12000, 40
    A = synth.snapshots(m, n, 77)
    R = np.zeros((n, n)); W = np.zeros((n, n)); S = np.zeros(n); V = np.zeros((n, n)); U = np.zeros((m, n))
    assert L.pl_tsqr_host_apply_f64(U.ctypes.data, W.ctypes.data, m, n) != 0          # nothing factored yet
    assert L.pl_tsqr_host_factor_f64(R.ctypes.data, A.ctypes.data, m, n) == 0, L.pl_last_error()
    assert np.allclose(np.tril(R, -1), 0)
    Sa = np.linalg.svd(A, compute_uv=False)
    assert np.abs(np.linalg.svd(R, compute_uv=False) - Sa).max() <= 1e-13 * Sa[0]     # R^T R = A^T A
    assert L.pl_tsqr_host_stack_f64(W.ctypes.data, S.ctypes.data, V.ctypes.data, R.ctypes.data, 1, n) == 0, L.pl_last_error()
    assert L.pl_tsqr_host_apply_f64(U.ctypes.data, W.ctypes.data, m + 1, n) != 0      # wrong size is refused ...
    assert L.pl_tsqr_host_apply_f64(U.ctypes.data, W.ctypes.data, m, n) == 0, L.pl_last_error()   # ... and the state survives
    assert_svd_parity(po.tsqr_svd(A), (U, S, V))
    assert L.pl_tsqr_host_apply_f64(U.ctypes.data, W.ctypes.data, m, n) != 0          # consumed
    L.pl_host_cache_free()


@pytest.mark.parametrize("m,a,b", [(5000, 16, 64), (5000, 64, 16), (40000, 8, 512), (33333, 33, 151), (20000, 151, 151),
                                    (3001, 24, 40), (70, 5, 9), (100000, 200, 96), (17, 64, 64), (250000, 12, 999)])
def test_matmul_tn_kernel(pl, m, a, b):
    """C = X^T Y (the rank-local part of matmulp) against torch's fp64 matmul: all three tile shapes, the swapped
    (a > b) orientation, odd leading dimensions (8-byte loads) and row counts that are not a multiple of the stage."""
    g = torch.Generator(device="cuda"); g.manual_seed(m + a)
    X = torch.randn((m, a), dtype=torch.float64, device="cuda", generator=g)
    Y = torch.randn((m, b), dtype=torch.float64, device="cuda", generator=g)
    from pyloworder_b200.vmmath.maths import matmul_tn
    C = matmul_tn(X, Y)
    ref = X.T @ Y
    tol = 1e-13 * m ** 0.5 * 8
    assert float((C - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    C2 = matmul_tn(X, Y)
    assert torch.equal(C, C2), "split-K reduction must be deterministic"
    # the public call shape: matmulp(Ai.T, Qi) on a .T view, and on strided column views
    C3 = pl.math.matmulp(X.T, Y)
    assert torch.equal(C3, C)
    if a >= 8 and b >= 8:
        C4 = pl.math.matmulp(X[:, 1:a - 2].T, Y[:, 3:b - 1])
        assert float((C4 - ref[1:a - 2, 3:b - 1]).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    Cn = pl.math.matmulp(host(X).T, host(Y))
    assert isinstance(Cn, np.ndarray) and np.abs(Cn - host(ref)).max() <= tol * max(1.0, float(ref.abs().max()))


def test_matmul_tn_degenerate(pl):
    """Empty contraction (a rank without rows) and 1 x 1 outputs."""
    from pyloworder_b200.vmmath.maths import matmul_tn
    X = torch.empty((0, 7), dtype=torch.float64, device="cuda"); Y = torch.empty((0, 5), dtype=torch.float64, device="cuda")
    C = matmul_tn(X, Y)
    assert C.shape == (7, 5) and float(C.abs().max()) == 0.0
    x = torch.arange(1.0, 1001.0, dtype=torch.float64, device="cuda").reshape(-1, 1)
    assert float(matmul_tn(x, x)[0, 0]) == float((x * x).sum())


def test_svd_rectangular(pl):
    """svd() of the wide (r x n) matrix B of randomized_svd and of a local tall matrix, against LAPACK."""
    rng = np.random.default_rng(3)
    for shape in [(8, 40), (12, 96), (1, 7), (64, 65), (300, 20)]:
        B = rng.standard_normal(shape)
        U, S, V = [host(t) for t in pl.math.svd(dev(B))]
        k = min(shape)
        assert U.shape == (shape[0], k) and S.shape == (k,) and V.shape == (k, shape[1])
        So = np.linalg.svd(B, compute_uv=False)
        assert np.abs(S - So).max() <= 1e-13 * So[0]
        assert np.abs((U * S) @ V - B).max() <= 1e-13 * So[0] * 4
        assert np.abs(U.T @ U - np.eye(k)).max() <= 1e-13 and np.abs(V @ V.T - np.eye(k)).max() <= 1e-13


@pytest.mark.parametrize("path", RSVD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_randomized_svd_against_reference_golden(pl, path):
    """randomized_qr / randomized_svd / POD.run(randomized=True) against the reference's own Python output for the
    same seed (the sketch matrix is bit-identical: numpy MT19937 on the host, as in pyLOM/vmmath/svd.py:131-133)."""
    g = np.load(path)
    A, r, q, sk = g["A"], int(g["r"]), int(g["q"]), int(g["seed"])
    Ad = dev(A)
    Q, B = [host(t) for t in pl.math.randomized_qr(Ad, r, q, seed=sk)]
    assert Q.shape == (A.shape[0], r) and B.shape == (r, A.shape[1])
    assert np.abs(Q.T @ Q - np.eye(r)).max() <= 1e-12
    assert np.abs(Q.T @ A - B).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
    assert np.abs(g["Q"] @ (g["Q"].T @ Q) - Q).max() <= 1e-9              # same range as the reference's Q
    U, S, V = [host(t) for t in pl.math.randomized_svd(Ad, r, q, seed=sk)]
    assert U.shape == (A.shape[0], r) and S.shape == (r,) and V.shape == (r, A.shape[1])
    assert_svd_parity((g["U"], g["S"], g["V"]), (U, S, V))
    Up, Sp, Vp = [host(t) for t in pl.POD.run(Ad, remove_mean=True, randomized=True, r=r, q=q, seed=sk)]
    assert torch.equal(Ad, dev(A)), "POD.run must not modify X"
    assert_svd_parity((g["pod_U"], g["pod_S"], g["pod_V"]), (Up, Sp, Vp))
    # numpy in -> numpy out
    Un, Sn, Vn = pl.math.randomized_svd(A, r, q, seed=sk)
    assert isinstance(Un, np.ndarray) and np.abs(Sn - S).max() <= 1e-12 * S[0]


@pytest.mark.parametrize("path", RSVD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_streaming_randomized_qr(pl, path):
    """init_qr_streaming / update_qr_streaming.  The first block is range-identical to the reference's; the update
    adds two sketches whose column signs come from the QR (result depends on the QR's sign convention, also between
    the reference's own P = 1 and P = 2 runs), so it is checked through the identities the algorithm guarantees and
    through the approximation error, which must be as good as the reference's."""
    g = np.load(path)
    A, r, q, sk, n1 = g["A"], int(g["r"]), int(g["q"]), int(g["seed"]), int(g["st_n1"])
    A1, A2 = np.ascontiguousarray(A[:, :n1]), np.ascontiguousarray(A[:, n1:])
    Q1, B1, Y1 = pl.math.init_qr_streaming(dev(A1), r, q, seed=sk)
    Q1h = host(Q1)
    assert np.abs(g["st_Q1"] @ (g["st_Q1"].T @ Q1h) - Q1h).max() <= 1e-9          # same range as the reference's Q1
    assert np.abs(Q1h.T @ A1 - host(B1)).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
    Q2, B2, Y2 = [host(t) for t in pl.math.update_qr_streaming(dev(A2), Q1, B1, Y1, r, q)]
    assert Q2.shape == (A.shape[0], r) and B2.shape == (r, A.shape[1]) and Y2.shape == (A.shape[0], r)
    assert np.abs(Q2.T @ Q2 - np.eye(r)).max() <= 1e-12
    assert np.abs(Q2 @ (Q2.T @ Y2) - Y2).max() <= 1e-10 * np.abs(Y2).max()
    assert np.abs(B2[:, n1:] - Q2.T @ A2).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
    assert np.abs(B2[:, :n1] - (Q2.T @ Q1h) @ host(B1)).max() <= 1e-12 * np.abs(A).max() * A.shape[1]
    err = np.linalg.norm(A - Q2 @ B2) / np.linalg.norm(A)
    err_ref = np.linalg.norm(A - g["st_Q2"] @ g["st_B2"]) / np.linalg.norm(A)
    assert err <= 2.0 * err_ref + 1e-12, (err, err_ref)


@pytest.mark.parametrize("m,n,r,q", [(200000, 128, 16, 2), (60000, 512, 40, 1), (30000, 151, 10, 3)])
def test_randomized_svd_against_oracle(pl, m, n, r, q):
    A = synth.snapshots(m, n, 31)
    ref = po.randomized_svd(A, r, q, 5)
    got = [host(t) for t in pl.math.randomized_svd(dev(A), r, q, seed=5)]
    assert_svd_parity(ref, got)
    # the leading modes approximate the deterministic ones
    S_full = np.linalg.svd(A, compute_uv=False)
    assert np.abs(got[1][:4] - S_full[:4]).max() <= 1e-6 * S_full[0]


@pytest.mark.parametrize("path", DMD_GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_dmd_against_reference_golden(pl, path):
    """DMD.run / frequency_damping / reconstruction_jovanovic against the reference's own Python output.  Modes are
    unique up to a unit complex factor per mode (LAPACK fixes only the norm of an eigenvector), so eigenvalues, |b|,
    the products Phi_k b_k and the reconstruction are compared; the amplitudes solve a Vandermonde-Gram system with
    cond ~ 1e8, hence 1e-6 there and 1e-10 on the eigenvalues."""
    g = np.load(path)
    X, r, dt = g["X"], float(g["r"]), float(g["dt"])
    Xd = dev(X)
    muR, muI, Phi, b = pl.DMD.run(Xd, r, remove_mean=True)
    assert torch.equal(Xd, dev(X)), "DMD.run must not modify X"
    assert Phi.dtype == torch.complex128 and Phi.shape == (X.shape[0], g["P1_muReal"].shape[0])
    muRh, muIh, Phih, bh = host(muR), host(muI), host(Phi), host(b)
    assert np.abs(muRh - g["P1_muReal"]).max() <= 1e-10 and np.abs(muIh - g["P1_muImag"]).max() <= 1e-10
    bref, Pref = g["P1_b"], g["P1_Phi"]
    assert np.abs(np.abs(bh) - np.abs(bref)).max() <= 1e-6 * np.abs(bref).max()
    assert np.abs(Phih * bh - Pref * bref).max() <= 1e-6 * np.abs(Pref * bref).max()
    assert np.all(muIh[0::2] >= 0)
    delta, omega = pl.DMD.frequency_damping(muR, muI, dt)
    assert np.abs(host(delta) - g["delta"]).max() <= 1e-9 and np.abs(host(omega) - g["omega"]).max() <= 1e-9
    t = np.arange(X.shape[1], dtype=np.double)
    Xr = host(pl.DMD.reconstruction_jovanovic(Phi, muR, muI, t, b))
    assert np.abs(Xr - g["X_DMD"]).max() <= 1e-6 * np.abs(g["X_DMD"]).max()
    # numpy in -> numpy out
    muRn, muIn, Phin, bn = pl.DMD.run(X, r, remove_mean=True)
    assert isinstance(Phin, np.ndarray) and Phin.dtype == np.complex128 and np.abs(muRn - muRh).max() <= 1e-12
    # mode_computation: X V^T S^-1 |W|
    rng = np.random.default_rng(0)
    V = rng.standard_normal((5, X.shape[1])); S = rng.random(5) + 0.5; W = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    Mc = host(pl.DMD.mode_computation(Xd, V, S, W))
    assert np.abs(Mc - po.dmd_mode_computation(X, V, S, W)).max() <= 1e-11 * np.abs(Mc).max()


def test_fused_input_read_matches(pl, monkeypatch):
    """n % 32 == 0 and m % 32 == 0 without centering: the first panel reads A directly instead of a padded copy.
    Same arithmetic, so the results must be bit-identical to the copy path, and A must stay untouched."""
    for (m, n) in ((64000, 96), (4096, 64), (300_000, 128)):
        A = dev(synth.snapshots(m, n, 13))
        A0 = A.clone()
        U1, S1, V1 = pl.math.tsqr_svd(A)
        assert torch.equal(A, A0)
        monkeypatch.setenv("PL_NO_FUSED_INPUT", "1")
        U0, S0, V0 = pl.math.tsqr_svd(A)
        monkeypatch.delenv("PL_NO_FUSED_INPUT")
        assert torch.equal(S0, S1) and torch.equal(V0, V1) and torch.equal(U0, U1)
        # poison the workspace: nothing of the factorisation buffer may be read before it is written
        from pyloworder_b200 import _dev
        for w in _dev._ws_cache.values():
            w.view(torch.float64)[: w.numel() // 8].fill_(float("nan")) if w.numel() % 8 == 0 else w.fill_(255)
        U2, S2, V2 = pl.math.tsqr_svd(A)
        assert torch.equal(S2, S1) and torch.equal(U2, U1)
        # the memory-saving in-place variant (output buffer = factorisation buffer) with the fused read
        monkeypatch.setenv("PL_INPLACE", "1")
        U3, S3, V3 = pl.math.tsqr_svd(A)
        monkeypatch.delenv("PL_INPLACE")
        assert torch.equal(A, A0) and torch.equal(S3, S1) and torch.equal(V3, V1)
        assert float((U3 - U1).abs().max()) <= 1e-14


def test_form_q_then_apply_matches(pl):
    """apply_q in two steps (flags bit 1: form Q only; then bit 0: multiply) -- the split the multi-rank path uses to
    overlap the Q formation with the exchange -- gives the same U as the one-step call, also on a side stream."""
    from pyloworder_b200.vmmath.svd import CudaEngine
    eng = CudaEngine()
    A = dev(synth.snapshots(50000, 72, 3))
    W = dev(np.linalg.qr(np.random.default_rng(1).standard_normal((72, 72)))[0])
    eng.factor(A, "t_one")
    U1 = eng.apply_q(A.shape, W, "t_one", A.device)
    R, _ = eng.factor(A, "t_two")
    eng.form_q(A.shape, "t_two", A.device)
    side = eng.side_stream(A.device)
    assert side is not None
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        W2 = W.clone()                       # stands for the small factorisations done on the side stream
        W2.record_stream(torch.cuda.current_stream(A.device))
    torch.cuda.current_stream().wait_stream(side)
    U2 = eng.apply_q(A.shape, W2, "t_two", A.device, formed=True)
    assert torch.equal(U1, U2)
    assert float((U2.T @ U2 - torch.eye(72, dtype=torch.float64, device="cuda")).abs().max()) <= 1e-13


def test_lookahead_schedule_matches(pl, monkeypatch):
    """The optional two-stream panel look-ahead (PL_LOOKAHEAD=1) reorders launches only: same R, same U."""
    m, n = 300_000, 96                      # >= 2048 tiles, the size from which the look-ahead engages
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    Q0, R0 = pl.math.qr(A)
    monkeypatch.setenv("PL_LOOKAHEAD", "1")
    Q1, R1 = pl.math.qr(A)
    assert torch.equal(R0, R1) and torch.equal(Q0, Q1)
    I = torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((Q1.T @ Q1 - I).abs().max()) <= 1e-13
    assert float((Q1 @ R1 - A).abs().max()) <= 1e-12


def test_large_properties(pl):
    """Size-independent properties at a size the CPU oracle would not finish quickly."""
    m, n = 2_000_000, 64
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    A[:, 5] = A[:, 4] * 2.0 + 1e-9 * A[:, 5]          # a nearly dependent column
    U, S, V = pl.math.tsqr_svd(A)
    I = torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((U.T @ U - I).abs().max()) <= 1e-12
    assert float((V @ V.T - I).abs().max()) <= 1e-12
    assert float(((U * S) @ V - A).norm() / A.norm()) <= 1e-13
    assert bool((S[:-1] >= S[1:]).all())
    # linearity: scaling A scales S, leaves the modes (up to sign)
    U2, S2, V2 = pl.math.tsqr_svd(A * 3.0)
    assert float((S2 - 3.0 * S).abs().max() / (3.0 * S[0])) <= 1e-13
    # idempotence of the projector on the range
    X = pl.POD.reconstruct(U, S, V)
    assert float((X - A).abs().max()) <= 1e-11


def test_inplace_variant_matches(pl, monkeypatch):
    """The memory-saving in-place path (output buffer = factorisation buffer) gives the same answer."""
    A = synth.snapshots(30000, 96, 4)
    monkeypatch.setenv("PL_NO_INPLACE", "1")
    U0, S0, V0 = [host(t) for t in pl.POD.run(dev(A), remove_mean=True)]
    monkeypatch.delenv("PL_NO_INPLACE"); monkeypatch.setenv("PL_INPLACE", "1")
    U1, S1, V1 = [host(t) for t in pl.POD.run(dev(A), remove_mean=True)]
    Q1, R1 = [host(t) for t in pl.math.qr(dev(A))]
    assert np.abs(S0 - S1).max() <= 1e-14 * S0[0]
    assert np.abs(U0 - U1).max() <= 1e-13 and np.abs(V0 - V1).max() <= 1e-13
    assert np.abs(Q1 @ R1 - A).max() <= 1e-12 * np.abs(A).max()


def test_native_library_is_the_one_running(pl):
    from pyloworder_b200 import _lib
    before = _lib.lib().pl_launch_count()
    pl.math.tsqr_svd(dev(synth.random_matrix(500, 10, 2)))
    assert _lib.lib().pl_launch_count() > before
    maps = open("/proc/self/maps").read()
    assert "libpylom_b200.so" in maps


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_parity():
    """One process per GPU over NCCL (tests/dist_check.py): tsqr_svd / POD.run vs the oracle on the same shards."""
    import socket, subprocess, sys
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    script = os.path.join(os.path.dirname(__file__), "dist_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c_abi_two_ranks_no_torch(tmp_path):
    """tests/c_abi_dist.c: two processes (one per GPU), no Python and no torch, through pl_get_unique_id /
    pl_comm_init_rank / pl_tsqr_svd_host_dist_f64 -- the one-call collective that replaces the reference's dtsqr_svd
    (pyLOM/vmmath/src/svd.c:678-712) on P ranks.  Compiled here with gcc against include/pylom_b200.h."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "pyloworder_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "c_abi_dist")
    cc = subprocess.run(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "c_abi_dist.c"), "-o", exe,
                         "-L" + libdir, "-lpylom_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                         "-Wl,-rpath," + libdir + ",-rpath," + os.path.join(cuda, "lib64")], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "C_ABI_DIST PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_shape_fuzz(pl):
    """Random (m, n) including tile / block / panel boundary cases: invariants against numpy."""
    rng = np.random.default_rng(123)
    shapes = [(32, 32), (33, 32), (127, 32), (128, 32), (129, 33), (160, 64), (161, 65), (4096, 96), (4097, 97),
              (512 * 13 + 5, 40), (65, 65), (1024, 1), (2, 2), (1, 1)]
    for _ in range(14):
        n = int(rng.integers(1, 200))
        m = int(n + rng.integers(0, 40000))
        shapes.append((m, n))
    for (m, n) in shapes:
        A = rng.standard_normal((m, n)) * np.exp(rng.uniform(-3, 3, size=(1, n)))
        U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(A))]
        So = np.linalg.svd(A, compute_uv=False)
        assert np.abs(S - So).max() <= 1e-12 * So[0], (m, n)
        assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-12, (m, n)
        assert np.abs(V @ V.T - np.eye(n)).max() <= 1e-12, (m, n)
        assert np.abs((U * S) @ V - A).max() <= 1e-11 * np.abs(A).max(), (m, n)
        mean = host(pl.math.temporal_mean(dev(A)))
        assert np.abs(mean - A.mean(1)).max() <= 1e-13 * np.abs(A).max()


# ---- small-n fused tile TSQR (csrc/tsqr_small.cu): n <= 64, the shape of BASELINE config 5 -----------------------
@pytest.fixture
def small_path(monkeypatch):
    """Force the small-n path on test-sized inputs: many short strips, partial last tile."""
    monkeypatch.setenv("PL_SMALL_MIN_ROWS", "2048")
    monkeypatch.setenv("PL_SMALL_MIN_TILES", "1")
    monkeypatch.setenv("PL_SMALL_STRIPS", "7")
    monkeypatch.delenv("PL_NO_SMALL", raising=False)


@pytest.mark.parametrize("m,n,kind", [(9000, 64, "rand"), (20077, 64, "def"), (12345, 50, "rand"), (8192, 32, "rand"),
                                      (9000, 17, "def"), (6000, 3, "rand"), (40000, 64, "cond")])
def test_small_path_qr(pl, small_path, m, n, kind):
    rng = np.random.default_rng(m + n)
    A = synth.random_matrix(m, n, 5, cond=1e11) if kind == "cond" else rng.standard_normal((m, n))
    if kind == "def":
        A[:, n // 3] = 0.0; A[:, n - 1] = A[:, 1]
    Q, R = [host(t) for t in pl.math.qr(dev(A))]
    sc = np.abs(A).max()
    assert np.allclose(np.tril(R, -1), 0)
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 1e-13
    assert np.abs(Q @ R - A).max() <= 1e-13 * sc * n
    if kind == "def":     # R is not unique below a zero pivot: only the invariants above apply
        return
    Rr = np.linalg.qr(A, mode="r")
    assert np.abs(np.abs(R) - np.abs(Rr)).max() <= 1e-11 * np.abs(Rr).max()
    # same matrix through the generic CAQR path: identical R up to row signs
    import os
    os.environ["PL_NO_SMALL"] = "1"
    try:
        Rg = host(pl.math.qr(dev(A))[1])
    finally:
        del os.environ["PL_NO_SMALL"]
    assert np.abs(np.abs(R) - np.abs(Rg)).max() <= 1e-11 * np.abs(Rr).max()


@pytest.mark.parametrize("m,n,inplace", [(20000, 64, False), (20000, 64, True), (33333, 40, False), (16448, 32, True)])
def test_small_path_pod_against_oracle(pl, small_path, monkeypatch, m, n, inplace):
    if inplace:
        monkeypatch.setenv("PL_INPLACE", "1")
    X = synth.snapshots(m, n, 11)
    U, S, V = pl.POD.run(dev(X), remove_mean=True)
    assert_svd_parity(po.pod_run(X, remove_mean=True), (host(U), host(S), host(V)))
    Y = X - X.mean(axis=1, keepdims=True)
    Uh, Sh, Vh = host(U), host(S), host(V)
    assert np.abs(Uh.T @ Uh - np.eye(n)).max() <= 1e-12
    assert np.abs((Uh * Sh) @ Vh - Y).max() <= 1e-12 * np.abs(Y).max()
    U2, S2, V2 = [host(t) for t in pl.math.tsqr_svd(dev(X))]
    assert_svd_parity(po.tsqr_svd(X), (U2, S2, V2))
    # variance normalisation goes through the reflector store (centre/scale pass + in-place factorisation)
    U3, S3, V3 = [host(t) for t in pl.POD.run(dev(X), remove_mean=True, divide_variance=True)]
    assert_svd_parity(po.pod_run(X, remove_mean=True, divide_variance=True), (U3, S3, V3))


def test_small_path_rank_deficient_svd(pl, small_path):
    """Two numerically zero singular values (a zero column and a duplicated one): the Jacobi noise floor ends the
    sweeps, U and V stay orthonormal."""
    rng = np.random.default_rng(3)
    for m, n in ((20077, 64), (9000, 17)):
        A = rng.standard_normal((m, n)); A[:, n // 3] = 0.0; A[:, n - 1] = A[:, 1]
        U, S, V = [host(t) for t in pl.math.tsqr_svd(dev(A))]
        So = np.linalg.svd(A, compute_uv=False)
        assert np.all(np.isfinite(S)) and np.abs(S - So).max() <= 1e-13 * So[0]
        assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-12
        assert np.abs(V @ V.T - np.eye(n)).max() <= 1e-12
        assert np.abs((U * S) @ V - A).max() <= 1e-12 * np.abs(A).max()
