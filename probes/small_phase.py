"""Phase cycle counters of small_factor_kernel (library built with EXTRA=-DPL_SMALL_TIMING)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
L = _lib.lib()
L.pl_debug_small_read.argtypes = [ctypes.c_void_p]
for (m, n) in [(16_000_000, 64), (16_000_000, 32)]:
    A = torch.rand((m, n), dtype=torch.float64, device="cuda")
    buf = (ctypes.c_ulonglong * 8)()
    pl.math.tsqr_svd(A); L.pl_debug_small_read(buf)
    pl.math.tsqr_svd(A); L.pl_debug_small_read(buf)
    tiles = m // 512
    names = ["load", "gemm1", "small", "gemm2", "chain", "store+gram+T", "head", "tstore"]
    tot = sum(buf)
    print(f"m={m} n={n}: cycles per tile per CTA: " + " ".join(f"{names[i]}={buf[i] / tiles:.0f}" for i in range(8)) + f" total={tot / tiles:.0f}", flush=True)
    del A
