"""Time pl_qr_factor_f64 with a variant library (argv[1]) on m x n."""
import sys, os, ctypes
import torch
L = ctypes.CDLL(sys.argv[1])
L.pl_qr_workspace_bytes.restype = ctypes.c_size_t
L.pl_qr_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
for (m, n) in ((8000000, 32), (1000000, 32)):
    A = torch.randn((m, n), dtype=torch.float64, device="cuda")
    R = torch.empty((n, n), dtype=torch.float64, device="cuda")
    wsb = L.pl_qr_workspace_bytes(m, n)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    wp = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    def run():
        rc = L.pl_qr_factor_f64(ctypes.c_void_p(R.data_ptr()), None, ctypes.c_void_p(A.data_ptr()), ctypes.c_int64(m), ctypes.c_int64(n),
                                0, ctypes.c_void_p(wp), ctypes.c_size_t(wsb), None)
        assert rc == 0
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    print(os.path.basename(sys.argv[1]), m, n, "factor ms", e0.elapsed_time(e1) / 2, flush=True)
    del A, ws
