"""Per-segment clock64 timing of the panel kernel (instrumented builds probes/libpl_timing_w*.so)."""
import sys, os, ctypes
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
for w in (1, 0, 4):
    L = ctypes.CDLL(os.path.join(root, "probes", f"libpl_timing_w{w}.so"))
    L.pl_qr_workspace_bytes.restype = ctypes.c_size_t
    L.pl_qr_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    for (m, n) in ((544, 32), (1_000_000, 32)):
        A = torch.randn((m, n), dtype=torch.float64, device="cuda")
        R = torch.empty((n, n), dtype=torch.float64, device="cuda")
        wsb = L.pl_qr_workspace_bytes(m, n)
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
        wp = ws.data_ptr() + ((-ws.data_ptr()) % 256)
        out = (ctypes.c_ulonglong * 8)()
        L.pl_debug_panel_read(out)
        rc = L.pl_qr_factor_f64(ctypes.c_void_p(R.data_ptr()), None, ctypes.c_void_p(A.data_ptr()), ctypes.c_int64(m), ctypes.c_int64(n),
                                0, ctypes.c_void_p(wp), ctypes.c_size_t(wsb), None)
        assert rc == 0
        L.pl_debug_panel_read(out)
        v = [int(x) for x in out]
        tot = sum(v[:6])
        names = ["loop/top", "publish+dot", "barrier", "reduce+scalars", "update", "T+syncwarp", "tile-end-sync"]
        print(f"warp {w} m={m}: total loop cycles {tot}", {names[i]: round(v[i] / max(tot, 1), 3) for i in range(7)}, "abs/step(544 case: 4 tiles*32)", [round(x / 128) for x in v[:7]] if m == 544 else "")
