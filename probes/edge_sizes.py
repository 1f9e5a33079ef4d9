"""Edge-size sweep: wide matrices (n > 1024), tiny row counts, m == n, single column."""
import sys, time, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import pyloworder_b200 as pl
import synth
for (m, n) in ((1, 1), (2, 1), (3, 2), (31, 31), (33, 32), (129, 128), (1500, 1500), (4000, 1025), (5000, 1536), (6000, 2048), (100000, 3)):
    A = synth.random_matrix(m, n, 1)
    t0 = time.time()
    try:
        U, S, V = [t.cpu().numpy() for t in pl.math.tsqr_svd(torch.from_numpy(A).cuda())]
        So = np.linalg.svd(A, compute_uv=False)
        print(f"{m}x{n}: sigma_rel={np.abs(S - So).max() / So[0]:.2e} orthU={np.abs(U.T @ U - np.eye(n)).max():.2e} "
              f"orthV={np.abs(V @ V.T - np.eye(n)).max():.2e} recon={np.abs((U * S) @ V - A).max():.2e} {time.time() - t0:.2f}s", flush=True)
    except Exception as e:
        print(f"{m}x{n}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
