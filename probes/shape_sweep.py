"""Time tsqr_svd and its phases for several shapes (and planner settings via env)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyloworder_b200 as pl
from pyloworder_b200 import _lib, _dev
L = _lib.lib()
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(8000000, 64), (4000000, 256), (1000000, 999)]
names = ["copy", "panel", "upd_f", "upd_q", "gemm", "svd", "misc"]
for (m, n) in shapes:
    A = torch.randn((m, n), dtype=torch.float64, device="cuda")
    for _ in range(2):
        U, S, V = pl.math.tsqr_svd(A)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        U, S, V = pl.math.tsqr_svd(A)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    L.pl_profile_enable(1); pl.math.tsqr_svd(A); torch.cuda.synchronize(); L.pl_profile_enable(0)
    msb = (ctypes.c_double * 7)(); cnt = (ctypes.c_int64 * 7)()
    L.pl_profile_read(ctypes.cast(msb, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), 7)
    fl = 4.0 * m * n * n
    print(f"{m}x{n}: {ms:.1f} ms  {fl / ms * 1e-9:.2f} TFLOP/s_alg ({fl / ms * 1e-9 / 35.46 * 100:.1f}% roofline)  " +
          " ".join(f"{names[i]}={msb[i]:.1f}" for i in range(7)), flush=True)
    del A, U, S, V
    _dev.free_workspaces(); torch.cuda.empty_cache()
