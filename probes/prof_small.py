"""Short driver for ncu: one tsqr_svd of rows x cols (argv) on random data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyloworder_b200 as pl
m, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
A = torch.randn((m, n), dtype=torch.float64, device="cuda")
for _ in range(reps):
    U, S, V = pl.math.tsqr_svd(A)
torch.cuda.synchronize()
print("done", float(S[0]))
