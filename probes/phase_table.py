"""Per-kernel-class time of one tsqr_svd for a list of shapes (profiling hooks of the library).
Usage: python probes/phase_table.py 2000000x999 24000000x256 ..."""
import ctypes, json, sys, torch
sys.path.insert(0, ".")
import pyloworder_b200 as pl
from pyloworder_b200 import _lib, _dev
L = _lib.lib()
names = ["copy_center", "panel", "update_factor", "update_formq", "gemm", "svd_small", "misc"]
for spec in sys.argv[1:]:
    m, n = [int(v) for v in spec.split("x")]
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    for _ in range(2):
        U, S, V = pl.math.tsqr_svd(A); del U
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); U, S, V = pl.math.tsqr_svd(A); e1.record(); torch.cuda.synchronize(); del U
    total = e0.elapsed_time(e1)
    L.pl_profile_enable(1)
    U, S, V = pl.math.tsqr_svd(A); torch.cuda.synchronize(); del U
    L.pl_profile_enable(0)
    ms = (ctypes.c_double * 7)(); cnt = (ctypes.c_int64 * 7)()
    L.pl_profile_read(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), 7)
    print(json.dumps({"shape": spec, "total_ms": round(total, 2), "tflops_alg": round(4.0 * m * n * n / total * 1e-9, 2),
                      **{names[i]: round(ms[i], 2) for i in range(7)}}), flush=True)
    del A; _dev.free_workspaces(); torch.cuda.empty_cache()
