"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyloworder_b200 as pl
torch.manual_seed(0)
for (m, n) in ((3000, 70), (700, 33), (20000, 64)):
    A = torch.randn((m, n), dtype=torch.float64, device="cuda")
    U, S, V = pl.POD.run(A, remove_mean=True)
    Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=5)
    X = pl.POD.reconstruct(Ur, Sr, Vr)
    Q, R = pl.math.qr(A)
    torch.cuda.synchronize()
    print(m, n, float((U.T @ U - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()), flush=True)
os.environ["PL_INPLACE"] = "1"
A = torch.randn((5000, 64), dtype=torch.float64, device="cuda")
U, S, V = pl.math.tsqr_svd(A)
torch.cuda.synchronize()
print("inplace ok", float(S[0]))
