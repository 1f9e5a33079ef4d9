"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host layer has
no CPU fallback, planner / partitioning logic, bench's device generator == oracle generator, and
the multi-rank composition of tsqr_svd over gloo (world_size 2) with an injected CPU engine."""
import ctypes, os, re, subprocess, sys, textwrap
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pyloworder_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pylom_b200.h")).read()
    declared = set(re.findall(r"\b(pl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = ctypes.CDLL(_lib.libpath())
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/pylom_b200.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    assert _lib.lib().pl_version() >= 100


def test_workspace_query_and_argument_errors_without_gpu():
    from pyloworder_b200 import _lib
    L = _lib.lib()
    assert L.pl_qr_workspace_bytes(8_000_000, 512) > 8_000_000 * 512 * 8
    assert L.pl_qr_workspace_bytes(0, 5) == 0
    # bad shapes are rejected before any CUDA call
    rc = L.pl_tsqr_svd_f64(None, None, None, None, 3, 5, None, 0, None)
    assert rc < 0 and b"m >= n" in L.pl_last_error()
    rc = L.pl_qr_factor_f64(None, None, None, 100, 10, 0, None, 0, None)
    assert rc < 0 and b"workspace" in L.pl_last_error()


def test_host_pipeline_row_partition(monkeypatch):
    """Row chunks of the host-pointer pipeline (pure host arithmetic): they cover the shard exactly, all but the last
    are equal whole 128-row tiles, none is shorter than 4n rows when there are several, and PL_HOST_CHUNKS is a target,
    not a promise (the ragged / too-short cases that a first version got wrong)."""
    from pyloworder_b200 import _lib
    L = _lib.lib()
    buf = (ctypes.c_int64 * 64)()
    rng = np.random.default_rng(0)
    cases = [(8_000_000, 512, None), (4000, 64, 64), (9001, 33, 7), (20000, 100, 5), (512, 512, 8), (1, 1, 3), (1_000_000, 512, None),
             (125_000_000, 64, None), (2_000_000, 999, None)]
    cases += [(int(rng.integers(n, 3_000_000)), n, int(rng.integers(1, 70))) for n in rng.integers(1, 600, size=200).tolist()]
    for m, n, req in cases:
        if req is None:
            monkeypatch.delenv("PL_HOST_CHUNKS", raising=False)
        else:
            monkeypatch.setenv("PL_HOST_CHUNKS", str(req))
        C = L.pl_host_chunk_rows(m, n, ctypes.cast(buf, ctypes.c_void_p), 64)
        rows = [int(buf[c]) for c in range(C)]
        assert 1 <= C <= 64 and sum(rows) == m, (m, n, req, rows)
        assert all(r >= n for r in rows), (m, n, req, rows)
        if C > 1:
            assert len(set(rows[:-1])) == 1 and rows[0] % 128 == 0 and min(rows) >= 4 * n, (m, n, req, rows)
        if req is not None:
            assert C <= max(req, 1)
    monkeypatch.delenv("PL_HOST_CHUNKS", raising=False)
    assert L.pl_host_chunk_rows(8_000_000, 512, ctypes.cast(buf, ctypes.c_void_p), 64) == 16      # ~2 GiB per chunk
    assert L.pl_host_chunk_rows(3, 5, ctypes.cast(buf, ctypes.c_void_p), 64) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import pyloworder_b200 as pl
    X = np.random.default_rng(0).standard_normal((64, 4))
    for fn in (lambda: pl.POD.run(X), lambda: pl.math.tsqr_svd(X), lambda: pl.math.temporal_mean(X)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_worksplit_matches_oracle():
    import pod_oracle as po
    from pyloworder_b200.utils import worksplit
    for m, P in ((10, 3), (89351, 8), (7, 7), (5, 8), (192_000_000, 8), (1000, 1)):
        for r in range(P):
            assert worksplit(0, m, r, P) == po.worksplit(0, m, r, P)


def test_truncation_rule_matches_golden(golden_dir):
    from pyloworder_b200.vmmath import compute_truncation_residual
    import glob
    for path in glob.glob(os.path.join(golden_dir, "*.npz")):
        if os.path.basename(path).startswith(("rsvd_", "dmd_")):
            continue
        g = np.load(path)
        S = g["tsqr_svd_P1_S"]
        for r, N in zip(g["trunc_r"], g["trunc_N"]):
            assert compute_truncation_residual(S, float(r)) == int(N)
            assert compute_truncation_residual(torch.from_numpy(S), float(r)) == int(N)


def test_complex_pair_selection_against_reference_golden(golden_dir):
    """Host half of the complex tsqr_svd (vmmath/svd.py:_complex_select): the SVD of the real embedding
    [[Ar, -Ai], [Ai, Ar]] (here by LAPACK, on the GPU by the Jacobi kernel) holds every complex singular triplet twice;
    the selection must return the reference's singular values and right vectors (up to a unit complex factor per mode)
    for the fixture made by the reference's own Python tsqr_svd, and for coinciding / zero singular values."""
    from pyloworder_b200.vmmath.svd import _complex_select
    g = np.load(os.path.join(golden_dir, "aux", "ztsqr_svd_500x12.npz"))
    rng = np.random.default_rng(3)
    n = 12
    Q, _ = np.linalg.qr(rng.standard_normal((200, n)) + 1j * rng.standard_normal((200, n)))
    W, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    sv = np.linspace(3.0, 1.0, n); sv[3] = sv[2]; sv[8] = sv[7] = sv[6]; sv[-2:] = 0.0
    for A, Sref, VHref in ((g["A"], g["P1_S"], g["P1_V"]), ((Q * sv) @ W.conj().T, np.sort(sv)[::-1], None)):
        Ahat = np.block([[A.real, -A.imag], [A.imag, A.real]])
        U2, S2, V2 = np.linalg.svd(Ahat, full_matrices=False)
        J, M, X = _complex_select(S2, V2, n)
        assert len(J) == n and len(set(J.tolist())) == n
        assert np.abs(S2[J] - Sref).max() <= 1e-13 * Sref[0]
        VH = X.conj().T
        assert np.abs(VH @ VH.conj().T - np.eye(n)).max() <= 1e-12
        Uc = (U2[:A.shape[0], J] + 1j * U2[A.shape[0]:, J])
        if M is not None:
            Uc = Uc @ M
        assert np.abs((Uc * S2[J]) @ VH - A).max() <= 1e-12 * np.abs(A).max() * n
        if VHref is not None:      # separated spectrum: rows of V^H equal the reference's up to a phase
            assert M is None
            assert np.abs(np.einsum("kj,kj->k", VHref.conj(), VH)).min() >= 1 - 1e-10


def test_division_free_jacobi_rotation_formula():
    """Executable spec of `jacobi_rotation` (csrc/svd_small.cu): t = sgn(d) h / (|d| + sqrt(d^2 + h^2)) with d = b - a,
    h = 2 c is the textbook rotation t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 c), including the
    sign conventions at d = +-0 and c < 0 (the kernel evaluates it with rsqrt / rcp seeds + Newton steps)."""
    def lib(a, b, c):
        d, h = b - a, 2.0 * c
        den = abs(d) + np.sqrt(d * d + h * h)
        t = (-h if np.signbit(d) else h) / den
        cs = 1.0 / np.sqrt(1.0 + t * t)
        return cs, cs * t
    def book(a, b, c):
        zeta = (b - a) / (2.0 * c)
        t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
        cs = 1.0 / np.sqrt(1.0 + t * t)
        return cs, cs * t
    rng = np.random.default_rng(0)
    cases = [(1.0, 1.0, 0.3), (1.0, 1.0, -0.3), (2.0, 1.0, 0.5), (1.0, 2.0, -0.5), (1e-8, 1e8, 1e-3), (1e8, 1e-8, -1e-3)]
    for _ in range(5000):
        a, b = rng.random(2) * 10.0 ** rng.uniform(-3, 3, 2)
        cases.append((a, b, rng.standard_normal() * np.sqrt(a * b) * rng.random()))
    for a, b, c in cases:
        if c == 0.0:
            continue
        (c1, s1), (c2, s2) = lib(a, b, c), book(a, b, c)
        assert abs(c1 - c2) <= 4e-16 and abs(s1 - s2) <= 4e-16, (a, b, c)
        # the rotation annihilates the inner product: rows x, y with |x|^2 = a, |y|^2 = b, x.y = c
        assert abs((c1 * c1 - s1 * s1) * c + c1 * s1 * (a - b)) <= 1e-14 * max(a, b)


def test_bench_generator_matches_oracle_generator():
    sys.path.insert(0, ROOT)
    import bench, synth
    X = bench.device_snapshots(torch, 5000, 24, 2022, 1000, 1400, torch.device("cpu"), chunk=150).numpy()
    Y = synth.snapshots(5000, 24, 2022, 1000, 1400)
    assert np.abs(X - Y).max() < 1e-13
    # the noise term (1e-8 * hash) must be the same hash, not just the smooth part
    assert np.abs((X - Y)).max() < 1e-8 * 1e-4


GLOO_SCRIPT = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'oracle'))
    import numpy as np, torch, torch.distributed as dist
    import pod_oracle as po, synth
    from pyloworder_b200.utils import parall
    import importlib; plsvd = importlib.import_module('pyloworder_b200.vmmath.svd')

    class CpuEngine:                      # test stand-in for the CUDA engine (numpy / LAPACK)
        def __init__(self): self.store = {{}}
        def factor(self, A, tag, center=False):
            a = A.numpy(); mean = None
            if center:
                mean = a.mean(1); a = a - mean[:, None]
            Q, R = np.linalg.qr(a); self.store[tag] = Q
            return torch.from_numpy(R), (torch.from_numpy(mean) if center else None)
        def apply_q(self, shape, W, tag, device):
            Q = self.store[tag]
            return torch.from_numpy(Q if W is None else Q @ W.numpy())
        def svd(self, R):
            U, S, V = np.linalg.svd(R.numpy()); return torch.from_numpy(U), torch.from_numpy(S), torch.from_numpy(V)
        def tsqr_svd_single(self, A, center=False): raise AssertionError('single-rank path in a 2-rank run')
        def allgather_rows(self, R): return parall.mpi_allgather_rows(R)
        def matmul(self, A, B): return torch.from_numpy(A.numpy() @ B.numpy())
        def matmul_tn(self, X, Y): return torch.from_numpy(X.numpy().T @ Y.numpy())
        def svd_any(self, A):
            U, S, V = np.linalg.svd(A.numpy(), full_matrices=False); return torch.from_numpy(U), torch.from_numpy(S), torch.from_numpy(V)

    rank, size = parall.init_distributed('gloo')
    assert size == 2 and parall.MPI_SIZE == 2
    m, n = 901, 12
    A = synth.snapshots(m, n, 5)
    r0, r1 = parall.worksplit(0, m, rank, size)
    U, S, V, mean = plsvd._tsqr_svd_dev(torch.from_numpy(A[r0:r1].copy()), center=True, engine=CpuEngine())
    Ul, So, Vo = po.pod_run([A[slice(*po.worksplit(0, m, r, 2))] for r in range(2)], remove_mean=True)
    mt = po.compare_svd(Ul[rank], So, Vo, U.numpy(), S.numpy(), V.numpy())
    assert mt['sigma_rel'] < 1e-13, mt
    assert mt['vmode_min'] > 1 - 1e-10, mt
    # U is distributed: inner products need the sum over ranks
    ip = np.einsum('ik,ik->k', Ul[rank], U.numpy())
    ip = parall.mpi_reduce(ip, op='sum')
    keep = So / So[0] > 1e-8
    assert np.abs(np.abs(ip[keep]) - 1).max() < 1e-8, ip
    Sg = [torch.zeros_like(S) for _ in range(2)]; dist.all_gather(Sg, S)
    assert torch.equal(Sg[0], Sg[1])            # identical on all ranks
    assert abs(float(parall.mpi_reduce(1.0)) - 2.0) < 1e-15
    # randomized path: sketch, power iterations (tsqr + matmulp = local X^T Y + all-reduce), rectangular svd
    Y = A - A.mean(1, keepdims=True)
    Ur_, Sr_, Vr_ = plsvd._randomized_svd_dev(torch.from_numpy(Y[r0:r1].copy()), 5, 2, 9, engine=CpuEngine())
    Ul, So, Vo = po.pod_run([A[slice(*po.worksplit(0, m, r, 2))] for r in range(2)], remove_mean=True, randomized=True, r=5, q=2, seed=9)
    assert Ur_.shape == (r1 - r0, 5) and Vr_.shape == (5, n)
    assert np.abs(Sr_.numpy() - So).max() < 1e-12 * So[0]
    ip = parall.mpi_reduce(np.einsum('ik,ik->k', Ul[rank], Ur_.numpy()), op='sum')
    assert np.abs(np.abs(ip) - 1).max() < 1e-8, ip
    # DMD on the POD basis: SVD of the first n-1 snapshots, projection U^T Y2 (local product + all-reduce), host eig
    from pyloworder_b200.DMD import wrapper as dmdw
    Xw = synth.dmd_waves(700, 24, 3)
    Yw = Xw - Xw.mean(1, keepdims=True)
    q0, q1 = parall.worksplit(0, 700, rank, size)
    muR, muI, Phi, bj = dmdw._run_dev(torch.from_numpy(Yw[q0:q1].copy()), 6, engine=CpuEngine())
    muRo, muIo, Phio, bo = po.dmd_run([Xw[slice(*po.worksplit(0, 700, r, 2))] for r in range(2)], 6)
    assert np.abs(muR - muRo).max() < 1e-10 and np.abs(muI - muIo).max() < 1e-10
    assert np.abs(Phi.numpy() * bj - Phio[rank] * bo).max() < 1e-6 * np.abs(bo).max()
    dist.barrier(); dist.destroy_process_group()
    print('RANK_OK_%d' % rank, flush=True)
""")


def test_two_rank_composition_over_gloo(tmp_path):
    script = tmp_path / "gloo2.py"
    script.write_text(GLOO_SCRIPT.format(root=ROOT))
    import socket
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0)); port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK_") == 2, r.stdout[-500:]
