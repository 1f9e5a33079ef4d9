"""tsqr_svd time for mid-size problems with the Jacobi/form-Q overlap forced on or off (one process per setting):
back-to-back calls, calls separated by a device synchronisation, and calls whose result is consumed (S read on the host)."""
import os, subprocess, sys
CHILD = r'''
import sys, time, torch
sys.path.insert(0, ".")
import pyloworder_b200 as pl
out = []
for (m, n) in ((89351, 151), (200000, 512), (500000, 512)):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    for _ in range(2): pl.math.tsqr_svd(A)
    torch.cuda.synchronize()
    res = []
    for mode in ("b2b", "sync", "read"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            U, S, V = pl.math.tsqr_svd(A)
            if mode == "sync": torch.cuda.synchronize()
            if mode == "read": s0 = float(S[0])
        torch.cuda.synchronize()
        res.append(f"{mode} {(time.perf_counter() - t0) / 5 * 1e3:.2f}")
    out.append(f"{m}x{n}: " + " ".join(res))
print(" | ".join(out))
'''
for label, env in (("overlap off", {"PL_NO_SVD_OVERLAP": "1"}), ("overlap on for all sizes", {"PL_SVD_OVERLAP_MIN_ROWS": "1"})):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True)
    print(f"{label:26s} {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]}", flush=True)
