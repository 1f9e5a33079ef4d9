"""Round-2 probe: where the time goes for the small shapes.  usage: python probes/r2_probe.py [cfg1] [svd] [small]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import pyloworder_b200 as pl
from pyloworder_b200 import _lib
import synth
L = _lib.lib()
NAMES = ["copy", "panel", "updF", "updQ", "gemm", "svd", "misc", "small"]


def classes(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    L.pl_profile_enable(1)
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 8)(); cnt = (ctypes.c_int64 * 8)()
    L.pl_profile_read(ms, cnt, 8); L.pl_profile_enable(0)
    return t, " ".join(f"{k}={ms[i] / reps:.3f}({cnt[i] // reps})" for i, k in enumerate(NAMES))


what = sys.argv[1:] or ["cfg1", "svd"]
if "cfg1" in what:
    X = torch.from_numpy(synth.snapshots(89351, 151, 2021)).cuda()
    t, c = classes(lambda: pl.POD.run(X, remove_mean=True))
    print(f"cfg1 POD.run 89351x151: {t:.3f} ms  {c}", flush=True)
    t, c = classes(lambda: pl.math.tsqr_svd(X))
    print(f"cfg1 tsqr_svd 89351x151: {t:.3f} ms  {c}", flush=True)
    os.environ["PL_DEBUG"] = "1"
    pl.POD.run(X, remove_mean=True); pl.math.tsqr_svd(X)
    os.environ.pop("PL_DEBUG")
if "svd" in what:
    for (m, n, seed, center) in [(200000, 512, 2022, False), (200000, 256, 2023, True), (100000, 999, 2024, False), (200000, 64, 2025, False), (89351, 151, 2021, True)]:
        A = synth.snapshots(m, n, seed)
        if center: A = A - A.mean(axis=1, keepdims=True)
        R = torch.from_numpy(np.linalg.qr(A, mode="r")).cuda().contiguous()
        G = torch.linalg.qr(torch.randn((4 * n, n), dtype=torch.float64, device="cuda"), mode="r")[1].contiguous()
        for name, M in (("synthetic", R), ("gaussian", G)):
            os.environ["PL_DEBUG"] = "1"
            pl.math.svd(M)
            os.environ.pop("PL_DEBUG")
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(3): U, S, V = pl.math.svd(M)
            torch.cuda.synchronize(); dt = (time.time() - t0) / 3
            So = np.linalg.svd(M.cpu().numpy(), compute_uv=False)
            err = np.abs(S.cpu().numpy() - So).max() / So[0]
            vv = (V @ V.T - torch.eye(n, device="cuda", dtype=torch.float64)).abs().max().item()
            uu = (U.T @ U - torch.eye(n, device="cuda", dtype=torch.float64)).abs().max().item()
            print(f"svd n={n} {name} center={center}: {dt * 1e3:.2f} ms  sigma err {err:.1e} VVt-I {vv:.1e} UtU-I {uu:.1e}", flush=True)
if "small" in what:
    L.pl_debug_small_read.argtypes = [ctypes.c_void_p]
    for (m, n) in [(32_000_000, 64)]:
        A = torch.rand((m, n), dtype=torch.float64, device="cuda")
        buf = (ctypes.c_ulonglong * 8)()
        pl.math.tsqr_svd(A); L.pl_debug_small_read(buf)
        pl.math.tsqr_svd(A); L.pl_debug_small_read(buf)
        tiles = m // 512
        names = ["load", "gemm1", "small", "gemm2", "chain", "store+gram+T", "head", "tstore"]
        tot = sum(buf)
        print(f"m={m} n={n}: cycles per tile per CTA: " + " ".join(f"{names[i]}={buf[i] / tiles:.0f}" for i in range(8)) + f" total={tot / tiles:.0f}", flush=True)
        del A
if "smalltime" in what:
    for (m, n) in [(32_000_000, 64), (125_000_000, 64), (32_000_000, 32)]:
        A = torch.rand((m, n), dtype=torch.float64, device="cuda")
        t, c = classes(lambda: pl.math.tsqr_svd(A))
        print(f"tsqr_svd {m}x{n}: {t:.3f} ms  {c}", flush=True)
        from pyloworder_b200.vmmath.svd import _engine
        W = torch.linalg.qr(torch.rand((n, n), dtype=torch.float64, device="cuda"))[0].contiguous()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = ta = 0.0
        for it in range(3):
            ev[0].record(); R, _ = _engine.factor(A, "local"); ev[1].record()
            U = _engine.apply_q((m, n), W, "local", A.device); ev[2].record(); torch.cuda.synchronize()
            if it: tf += ev[0].elapsed_time(ev[1]) / 2; ta += ev[1].elapsed_time(ev[2]) / 2
            del U
        print(f"   split: factor {tf:.2f} ms, apply {ta:.2f} ms", flush=True)
        del A; torch.cuda.empty_cache()
