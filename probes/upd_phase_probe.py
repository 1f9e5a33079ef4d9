"""Where does the update kernel's time go?  Runs tsqr_svd at rows x cols (pl.math.qr = factor + explicit Q) with the PL_UPD_DBG experiment flags
(results are garbage with any flag set; only the per-class timings are read).  One process per flag value.
Needs a library built with the hooks:  make -C pyloworder_b200/csrc clean && make -C pyloworder_b200/csrc EXTRA=-DPL_UPD_EXPERIMENTS"""
import ctypes, json, os, subprocess, sys
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
COLS = int(sys.argv[2]) if len(sys.argv) > 2 else 512
CHILD = r'''
import ctypes, json, sys, torch
sys.path.insert(0, ".")
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
L = _lib.lib()
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn((%d, %d), dtype=torch.float64, device="cuda", generator=g)
for _ in range(2):
    pl.math.qr(A)
torch.cuda.synchronize()
L.pl_profile_enable(1)
pl.math.qr(A); torch.cuda.synchronize()
L.pl_profile_enable(0)
ms = (ctypes.c_double * 7)(); cnt = (ctypes.c_int64 * 7)()
L.pl_profile_read(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), 7)
print(json.dumps({"update_factor": round(ms[2], 2), "update_formq": round(ms[3], 2), "panel": round(ms[1], 2)}))
''' % (ROWS, COLS)
ALL = ((0, "full"), (1, "no staging"), (2, "no GEMM1"), (4, "no T step"), (8, "no GEMM2 (no stores)"), (16, "no stores"),
                     (30, "staging only"), (1 | 16, "compute only, no stores"), (1 | 16 | 4 | 8, "GEMM1 only"),
                     (1 | 16 | 2 | 8, "T step only"), (1 | 16 | 2 | 4, "GEMM2 only"), (1 | 16 | 4, "GEMM1 + GEMM2"), (31, "barriers only"),
                     (64, "full, no L2 prefetch"), (64 | 1 | 16, "compute only, no stores, no prefetch"), (64 | 31, "barriers only, no prefetch"),
                     (64 | 1 | 16 | 4 | 8, "GEMM1 only, no prefetch"), (64 | 1 | 16 | 2 | 4, "GEMM2 only, no prefetch"), (64 | 1 | 16 | 4, "GEMM1 + GEMM2, no prefetch"),
                     (32, "full, GEMM1 fragments loaded once"), (32 | 1 | 16 | 4 | 8, "GEMM1 only, fragments loaded once"))
ONLY = [int(v) for v in os.environ.get("UPD_PROBE_ONLY", "").split(",") if v]
for flags, label in [fl for fl in ALL if not ONLY or fl[0] in ONLY]:
    env = dict(os.environ, PL_UPD_DBG=str(flags), PL_NO_SVD_OVERLAP="1")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"dbg={flags:2d} {label:28s} {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]}", flush=True)
