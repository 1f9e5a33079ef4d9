import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch, pyloworder_b200 as pl, synth
m, n = 12000, 999
A = torch.from_numpy(synth.snapshots(m, n, 2021)).cuda()
I = torch.eye(n, dtype=torch.float64, device="cuda")
Q, R = pl.math.qr(A)
print("Q orth", float((Q.T @ Q - I).abs().max()), "QR-A", float((Q @ R - A).abs().max()))
Ur, S, V = pl.math.svd(R)
print("Ur orth", float((Ur.T @ Ur - I).abs().max()), "V orth", float((V @ V.T - I).abs().max()), "recon", float(((Ur * S) @ V - R).abs().max()))
U = Q @ Ur
print("U=Q@Ur (torch) orth", float((U.T @ U - I).abs().max()))
U2, S2, V2 = pl.math.tsqr_svd(A)
print("tsqr_svd U orth", float((U2.T @ U2 - I).abs().max()))
d = (Ur.T @ Ur - I).abs()
print("Ur diag max", float(d.diagonal().max()), "offdiag max", float((d - torch.diag(d.diagonal())).max()))
