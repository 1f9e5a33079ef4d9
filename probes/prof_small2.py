import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyloworder_b200 as pl
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 64
A = torch.rand((m, n), dtype=torch.float64, device="cuda")
for _ in range(2):
    U, S, V = pl.math.tsqr_svd(A)
torch.cuda.synchronize()
