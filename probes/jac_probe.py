import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pyloworder_b200 as pl
for (m, n) in [(20077, 64), (9000, 17), (4000, 40)]:
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n)); A[:, n // 3] = 0.0; A[:, n - 1] = A[:, 1]
    U, S, V = pl.math.tsqr_svd(torch.from_numpy(A).cuda())
    print(m, n, "S tail", S[-3:].cpu().numpy(), flush=True)
