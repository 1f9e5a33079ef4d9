"""Timing probe for the randomized path: matmul_tn (X^T Y) against its HBM / FP64 bound and randomized_svd
end to end.  Usage: python probes/rsvd_time.py [rows] [cols]"""
import json, sys, time
import torch
sys.path.insert(0, ".")
import pyloworder_b200 as pl
from pyloworder_b200.vmmath.maths import matmul_tn

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
out = {"m": m, "n": n, "matmul_tn": [], "randomized_svd": []}


def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for a in (8, 16, 32, 64, 128, n):
    X = torch.randn((m, a), dtype=torch.float64, device="cuda", generator=g)
    ms = timeit(lambda: matmul_tn(X, A))
    byts = 8.0 * m * (a + n)
    fl = 2.0 * m * a * n
    out["matmul_tn"].append({"a": a, "b": n, "ms": round(ms, 3), "GBps": round(byts / ms * 1e-6, 1), "TFLOPs": round(fl / ms * 1e-9, 2)})
    print(out["matmul_tn"][-1], flush=True)
    del X
for r, q in ((16, 0), (16, 2), (64, 2)):
    ms = timeit(lambda: pl.math.randomized_svd(A, r, q, seed=1), reps=3)
    t0 = timeit(lambda: pl.math.matmul(A, torch.empty((n, r), dtype=torch.float64, device="cuda").normal_()), reps=3)
    out["randomized_svd"].append({"r": r, "q": q, "ms": round(ms, 3), "passes_over_A": 2 + 2 * q,
                                  "GBps_over_A": round((2 + 2 * q) * 8.0 * m * n / ms * 1e-6, 1), "matmul_A_omega_ms": round(t0, 3)})
    print(out["randomized_svd"][-1], flush=True)
ms = timeit(lambda: pl.math.tsqr_svd(A), reps=2)
out["tsqr_svd_ms"] = round(ms, 3)
print(json.dumps(out))
