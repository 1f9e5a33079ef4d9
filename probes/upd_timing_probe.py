"""Per-phase cycle shares of one warp of caqr_update_kernel (library built with EXTRA=-DPL_UPD_TIMING=<warp>)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
L = _lib.lib()
rd = ctypes.CDLL(_lib.libpath()).pl_debug_update_read
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn((m, 512), dtype=torch.float64, device="cuda", generator=g)
pl.math.qr(A); torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 8)(); rd(buf)
pl.math.qr(A); torch.cuda.synchronize()
rd(buf)
v = [int(x) for x in buf]; tot = sum(v)
names = ["issue staging", "wait data + barrier", "GEMM1 + epilogue", "barrier", "T step", "barrier", "GEMM2 + stores", "barrier + loop"]
print(" ".join(f"{n}: {100.0 * x / tot:.1f}%" for n, x in zip(names, v)))
