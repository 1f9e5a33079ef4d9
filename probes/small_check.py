"""Small-n fused TSQR path (csrc/tsqr_small.cu): correctness sweep + timings.  Usage: python probes/small_check.py [time]"""
import os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
import synth


def check(m, n, kind, inplace=False):
    rng = np.random.default_rng(m + n)
    if kind == "rand":
        A = rng.standard_normal((m, n))
    elif kind == "center":
        A = synth.snapshots(m, n, 3)
    else:
        A = rng.standard_normal((m, n)); A[:, n // 3] = 0.0; A[:, n - 1] = A[:, 1]
    os.environ.pop("PL_INPLACE", None)
    if inplace:
        os.environ["PL_INPLACE"] = "1"
    Ad = torch.from_numpy(A).cuda()
    t0 = time.time()
    if kind == "center":
        U, S, V = pl.POD.run(Ad, remove_mean=True)
        A = A - A.mean(axis=1, keepdims=True)
    else:
        U, S, V = pl.math.tsqr_svd(Ad)
    torch.cuda.synchronize()
    U, S, V = U.cpu().numpy(), S.cpu().numpy(), V.cpu().numpy()
    So = np.linalg.svd(A, compute_uv=False)
    e_s = np.abs(S - So).max() / So[0]
    e_o = np.abs(U.T @ U - np.eye(n)).max()
    e_r = np.abs((U * S) @ V - A).max() / np.abs(A).max()
    Q, R = pl.math.qr(Ad) if kind != "center" else (None, None)
    e_q = 0.0
    if Q is not None:
        Q, R = Q.cpu().numpy(), R.cpu().numpy()
        e_q = max(np.abs(Q.T @ Q - np.eye(n)).max(), np.abs(Q @ R - A).max() / np.abs(A).max())
    ok = e_s < 1e-12 and e_o < 1e-12 and e_r < 1e-12 and e_q < 1e-12
    print(f"{'OK ' if ok else 'BAD'} m={m} n={n} {kind} inplace={inplace}: sigma {e_s:.2e} orth {e_o:.2e} recon {e_r:.2e} qr {e_q:.2e}", flush=True)
    return ok


def timing(m, n, reps=3):
    Ad = torch.rand((m, n), dtype=torch.float64, device="cuda")
    L = _lib.lib()
    for _ in range(2):
        U, S, V = pl.math.tsqr_svd(Ad)
    torch.cuda.synchronize()
    L.pl_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        U, S, V = pl.math.tsqr_svd(Ad)
    e1.record(); torch.cuda.synchronize()
    import ctypes
    ms = (ctypes.c_double * 8)(); cnt = (ctypes.c_int64 * 8)()
    L.pl_profile_read(ms, cnt, 8)
    L.pl_profile_enable(0)
    t = e0.elapsed_time(e1) / reps
    roof = max(4.0 * m * n * n / 35.46e12, 32.0 * m * n / 6553.3e9) * 1e3
    print(f"time m={m} n={n}: {t:.2f} ms/step, roofline {roof:.2f} ms -> {roof / t:.3f}; classes(ms/step) "
          + " ".join(f"{k}={ms[i] / reps:.2f}" for i, k in enumerate(["copy", "panel", "updF", "updQ", "gemm", "svd", "misc", "small"])), flush=True)
    del U, Ad
    torch.cuda.empty_cache()


if __name__ == "__main__":
    allok = True
    for (m, n, kind) in [(8192, 64, "rand"), (20000, 64, "rand"), (20077, 64, "def"), (30011, 50, "rand"), (16384, 32, "rand"),
                         (9000, 17, "def"), (40000, 64, "center"), (33333, 40, "center"), (100000, 64, "rand"), (50000, 8, "rand")]:
        allok &= check(m, n, kind)
    for (m, n, kind) in [(20000, 64, "rand"), (40064, 32, "center")]:
        allok &= check(m, n, kind, inplace=True)
    os.environ.pop("PL_INPLACE", None)
    print("ALL OK" if allok else "FAILURES", flush=True)
    if len(sys.argv) > 1:
        for (m, n) in [(4_000_000, 64), (16_000_000, 64), (16_000_000, 32), (60_000_000, 64)]:
            timing(m, n)


def split_timing(m, n, reps=3):
    """pass 1 (+ stack factor) and pass 2 (+ stack apply) timed separately through the phase API."""
    from pyloworder_b200.vmmath.svd import _engine
    Ad = torch.rand((m, n), dtype=torch.float64, device="cuda")
    W = torch.linalg.qr(torch.rand((n, n), dtype=torch.float64, device="cuda"))[0].contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = ta = 0.0
    for it in range(reps + 1):
        ev[0].record()
        R, _ = _engine.factor(Ad, "local")
        ev[1].record()
        U = _engine.apply_q((m, n), W, "local", Ad.device)
        ev[2].record(); torch.cuda.synchronize()
        if it:
            tf += ev[0].elapsed_time(ev[1]); ta += ev[1].elapsed_time(ev[2])
        del U
    print(f"split m={m} n={n}: factor {tf / reps:.2f} ms, apply {ta / reps:.2f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "split":
    split_timing(16_000_000, 64)
    split_timing(16_000_000, 32)
