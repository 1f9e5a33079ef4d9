"""e2e through pl_tsqr_svd_host_f64 on plain numpy (pageable) arrays, with and without the library's pinned staging ring."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyloworder_b200 import _lib
L = _lib.lib()
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 512
rng = np.random.default_rng(0)
A = rng.random((m, n)); U = np.empty_like(A); S = np.empty(n); V = np.empty((n, n))
def call():
    _lib.check(L.pl_tsqr_svd_host_f64(U.ctypes.data, S.ctypes.data, V.ctypes.data, A.ctypes.data, m, n), "host")
call()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
t = min(ts)
print(f"pageable {m}x{n} staging={'off' if os.environ.get('PL_HOST_NO_STAGING') else 'on'}: {t * 1e3:.1f} ms = {4.0 * m * n * n / t * 1e-12:.2f} TFLOP/s, "
      f"{2 * m * n * 8 / t * 1e-9:.1f} GB/s over the link; UtU-I {np.abs(U[:, :8].T @ U[:, :8] - np.eye(8)).max():.1e} recon {np.abs((U[:1000] * S) @ V - A[:1000]).max():.1e}", flush=True)
