"""compute-sanitizer driver for the kernels and paths added after the first sanitizer pass: transposed-tall GEMM
(all tile shapes, aligned and 8-byte paths), narrow gemm_tall tiles, randomized SVD, DMD, the fused input read,
the chunked host pipeline and its phase calls."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pyloworder_b200 as pl
from pyloworder_b200 import _lib
from pyloworder_b200.vmmath.maths import matmul_tn
torch.manual_seed(0)
dev = "cuda"
for (m, a, b) in ((3000, 8, 70), (3000, 24, 64), (2500, 70, 33), (999, 151, 151)):
    X = torch.randn((m, a), dtype=torch.float64, device=dev); Y = torch.randn((m, b), dtype=torch.float64, device=dev)
    C = matmul_tn(X, Y); torch.cuda.synchronize()
    print("matmul_tn", m, a, b, float((C - X.T @ Y).abs().max()), flush=True)
A = torch.randn((6000, 96), dtype=torch.float64, device=dev)
for r in (5, 20, 40):
    U, S, V = pl.math.randomized_svd(A, r, 1, seed=3); torch.cuda.synchronize()
    print("randomized_svd", r, float(S[0]), flush=True)
Q, B, Yq = pl.math.init_qr_streaming(A[:, :48].contiguous(), 6, 1, seed=1)
Q2, B2, Y2 = pl.math.update_qr_streaming(A[:, 48:].contiguous(), Q, B, Yq, 6, 1); torch.cuda.synchronize()
print("streaming ok", tuple(B2.shape), flush=True)
t = torch.arange(40, dtype=torch.float64, device=dev) * 0.1
x = torch.linspace(0, 1, 4000, dtype=torch.float64, device=dev)[:, None]
Xd = torch.cos(6.28 * x - 2.0 * t) * torch.exp(-0.1 * t) + 0.5 * torch.cos(12.6 * x - 3.7 * t) + 1e-6 * torch.randn((4000, 40), dtype=torch.float64, device=dev)
muR, muI, Phi, bj = pl.DMD.run(Xd, 4, remove_mean=False)
Xr = pl.DMD.reconstruction_jovanovic(Phi, muR, muI, np.arange(40, dtype=float), bj); torch.cuda.synchronize()
print("DMD ok", float((Xr - Xd).abs().max()), flush=True)
A2 = torch.randn((4096, 64), dtype=torch.float64, device=dev)          # fused input read (n % 32 == 0, m % 32 == 0)
U, S, V = pl.math.tsqr_svd(A2); torch.cuda.synchronize()
print("fused read ok", float((U.T @ U - torch.eye(64, dtype=torch.float64, device=dev)).abs().max()), flush=True)
L = _lib.lib()
os.environ["PL_HOST_CHUNKS"] = "3"
for (m, n) in ((6000, 64), (5000, 50)):
    Ah = np.random.default_rng(0).standard_normal((m, n))
    Uh = np.zeros((m, n)); Sh = np.zeros(n); Vh = np.zeros((n, n)); R = np.zeros((n, n)); W = np.zeros((n, n))
    assert L.pl_tsqr_svd_host_f64(Uh.ctypes.data, Sh.ctypes.data, Vh.ctypes.data, Ah.ctypes.data, m, n) == 0
    assert L.pl_tsqr_host_factor_f64(R.ctypes.data, Ah.ctypes.data, m, n) == 0
    assert L.pl_tsqr_host_stack_f64(W.ctypes.data, Sh.ctypes.data, Vh.ctypes.data, R.ctypes.data, 1, n) == 0
    assert L.pl_tsqr_host_apply_f64(Uh.ctypes.data, W.ctypes.data, m, n) == 0
    print("host pipeline ok", m, n, float(np.abs(Uh.T @ Uh - np.eye(n)).max()), flush=True)
L.pl_host_cache_free()
print("done")
