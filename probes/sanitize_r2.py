"""compute-sanitizer driver for the round-2 kernels: second-generation update kernel (all template variants: factor /
form-Q, virtual-zero, 4- and 8-warp CTAs, upper tree levels, odd chunk counts, partial tiles), the small-n fused TSQR
and the SVD completion kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch, pyloworder_b200 as pl
import synth
torch.manual_seed(0)
dev = "cuda"
os.environ["PL_SMALL_MIN_ROWS"] = "2048"
for (m, n) in ((4096, 96), (5000, 100), (3001, 33), (20000, 160), (7777, 64), (9000, 40), (40000, 64)):
    A = torch.randn((m, n), dtype=torch.float64, device=dev)
    U, S, V = pl.math.tsqr_svd(A); torch.cuda.synchronize()
    I = torch.eye(n, dtype=torch.float64, device=dev)
    print("tsqr_svd", m, n, float((U.T @ U - I).abs().max()), float(((U * S) @ V - A).abs().max()), flush=True)
    Q, R = pl.math.qr(A); torch.cuda.synchronize()
    print("qr", m, n, float((Q @ R - A).abs().max()), flush=True)
X = torch.from_numpy(synth.snapshots(6000, 48, 7)).cuda()
U, S, V = pl.POD.run(X, remove_mean=True); torch.cuda.synchronize()
I = torch.eye(48, dtype=torch.float64, device=dev)
print("POD centred: VVt-I", float((V @ V.T - I).abs().max()), flush=True)
# second-generation Jacobi kernel (phantom rows, odd block counts), fp32 widening, complex embedding
for n in (65, 100, 151, 250):
    R = torch.linalg.qr(torch.randn((3 * n, n), dtype=torch.float64, device=dev), mode="r")[1].contiguous()
    U, S, V = pl.math.svd(R); torch.cuda.synchronize()
    print("svd", n, float(((U * S) @ V - R).abs().max()), flush=True)
X32 = torch.randn((5000, 40), dtype=torch.float32, device=dev)
U, S, V = pl.POD.run(X32, remove_mean=True); torch.cuda.synchronize()
print("fp32", U.dtype, float(S[0]), flush=True)
Ac = torch.randn((3000, 10), dtype=torch.complex128, device=dev)
U, S, VH = pl.math.tsqr_svd(Ac); torch.cuda.synchronize()
print("complex", float(((U * S) @ VH - Ac).abs().max()), flush=True)
