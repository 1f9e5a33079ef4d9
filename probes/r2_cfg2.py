"""cfg2-like timing with per-class breakdown: python probes/r2_cfg2.py [rows] [cols]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
L = _lib.lib()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
A = torch.rand((m, n), dtype=torch.float64, device="cuda")
NAMES = ["copy", "panel", "updF", "updQ", "gemm", "svd", "misc", "small"]
for _ in range(2): pl.math.tsqr_svd(A)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): U, S, V = pl.math.tsqr_svd(A)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3
L.pl_profile_enable(1)
U, S, V = pl.math.tsqr_svd(A); torch.cuda.synchronize()
ms = (ctypes.c_double * 8)(); cnt = (ctypes.c_int64 * 8)()
L.pl_profile_read(ms, cnt, 8); L.pl_profile_enable(0)
npad = -(-n // 32) * 32; Kp = npad // 32
fl = sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * (p + 1)) for p in range(Kp)) + sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * p) for p in range(Kp))
print(f"{m}x{n}: {t:.2f} ms/step = {4.0 * m * n * n / t * 1e-9:.2f} TF alg ({4.0 * m * n * n / t * 1e-9 / 35.46:.3f} of roofline); "
      + " ".join(f"{k}={ms[i]:.2f}({cnt[i]})" for i, k in enumerate(NAMES)) + f"; update kernels {fl / ((ms[2] + ms[3]) * 1e-3) * 1e-12:.2f} TF", flush=True)
I = torch.eye(n, dtype=torch.float64, device="cuda")
print("UtU-I", (U.T @ U - I).abs().max().item(), "VVt-I", (V @ V.T - I).abs().max().item(),
      "recon", ((U[:100000] * S) @ V - A[:100000]).abs().max().item())
