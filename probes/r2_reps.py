"""per-repetition timing of tsqr_svd: python probes/r2_reps.py rows cols [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pyloworder_b200 as pl
m, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
A = torch.rand((m, n), dtype=torch.float64, device="cuda")
for _ in range(2): pl.math.tsqr_svd(A)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    pl.math.tsqr_svd(A)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PL_"))
print(f"{m}x{n} [{tag}]: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
