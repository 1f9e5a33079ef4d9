"""clock64 phase attribution of caqr_panel_kernel (library built with -DPL_PANEL_TIMING -DPT_WARP=<warp>): python probes/r2_panel_timing.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
L = _lib.lib()
L.pl_debug_panel_read.argtypes = [ctypes.c_void_p]
names = ["top", "publish+dot", "barrier", "reduce+scalars", "update", "T+syncwarp"]
for (m, n) in ((4_000_000, 32), (4_000_000, 256)):
    A = torch.randn((m, n), dtype=torch.float64, device="cuda")
    out = (ctypes.c_ulonglong * 8)()
    pl.math.qr(A); L.pl_debug_panel_read(out)
    pl.math.qr(A); L.pl_debug_panel_read(out)
    v = [int(x) for x in out]
    steps = (m / 128) * 32 * (n // 32)          # level-0 column steps (upper levels add ~1 %)
    print(f"{m}x{n}: cycles per column step of one CTA:", {names[i]: round(v[i] / steps) for i in range(6)}, "sum", round(sum(v[:6]) / steps), flush=True)
