"""Stage-by-stage numerical check of the CUDA path against numpy (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import pyloworder_b200 as pl
import pod_oracle as po, synth

torch.cuda.init()
def dev(x): return torch.from_numpy(x).cuda()
def rep(name, err, tol):
    print(f"{'OK ' if err <= tol else 'BAD'} {name}: {err:.3e} (tol {tol:.1e})", flush=True)

# centering
for (m, n) in ((1000, 37), (513, 64), (300, 8), (200, 3), (4000, 512)):
    X = synth.snapshots(m, n, 1)
    mean = pl.math.temporal_mean(dev(X)).cpu().numpy()
    rep(f"temporal_mean {m}x{n}", np.abs(mean - X.mean(1)).max(), 1e-14)
    Y = pl.math.subtract_mean(dev(X), dev(mean)).cpu().numpy()
    rep(f"subtract_mean {m}x{n}", np.abs(Y - (X - mean[:, None])).max(), 0)
# gemm
for (m, n, k) in ((1000, 64, 64), (777, 151, 151), (5000, 512, 512), (300, 40, 7), (129, 33, 100)):
    A = np.random.default_rng(0).standard_normal((m, k)); B = np.random.default_rng(1).standard_normal((k, n))
    C = pl.math.matmul(dev(A), dev(B)).cpu().numpy()
    rep(f"matmul {m}x{n}x{k}", np.abs(C - A @ B).max(), 1e-12 * k)
# svd small
for n in (2, 5, 32, 77, 151):
    R = np.triu(np.random.default_rng(n).standard_normal((n, n)))
    U, S, V = [t.cpu().numpy() for t in pl.math.svd(dev(R))]
    Sref = np.linalg.svd(R, compute_uv=False)
    rep(f"svd n={n} sigma", np.abs(S - Sref).max() / Sref[0], 1e-14)
    rep(f"svd n={n} recon", np.abs((U * S) @ V - R).max(), 1e-13 * Sref[0])
    rep(f"svd n={n} orthU", np.abs(U.T @ U - np.eye(n)).max(), 1e-13)
# qr
for (m, n) in ((40, 40), (128, 32), (129, 32), (700, 24), (5000, 70), (20000, 8), (100000, 64), (30000, 151), (300000, 33)):
    A = synth.random_matrix(m, n, 4)
    t0 = time.time()
    Q, R = [t.cpu().numpy() for t in pl.math.qr(dev(A))]
    Rref = np.linalg.qr(A, mode="r")
    rep(f"qr {m}x{n} |R|", np.abs(np.abs(R) - np.abs(Rref)).max(), 1e-12 * n)
    rep(f"qr {m}x{n} orth", np.abs(Q.T @ Q - np.eye(n)).max(), 1e-13)
    rep(f"qr {m}x{n} QR-A", np.abs(Q @ R - A).max(), 1e-12)
# tsqr_svd
for (m, n, kind) in ((700, 24, "synth"), (30000, 151, "synth"), (100000, 64, "cond")):
    A = synth.snapshots(m, n, 2021) if kind == "synth" else synth.random_matrix(m, n, 3, cond=1e9)
    U, S, V = [t.cpu().numpy() for t in pl.math.tsqr_svd(dev(A))]
    Uo, So, Vo = po.tsqr_svd(A)
    mt = po.compare_svd(Uo, So, Vo, U, S, V)
    print(kind, m, n, mt, flush=True)
    rep(f"tsqr_svd {m}x{n} orth", np.abs(U.T @ U - np.eye(n)).max(), 1e-12)
    rep(f"tsqr_svd {m}x{n} recon", np.abs((U * S) @ V - A).max() / np.abs(A).max(), 1e-12)
# POD
X = synth.snapshots(20000, 40, 7)
U, S, V = pl.POD.run(dev(X), remove_mean=True)
Uo, So, Vo = po.pod_run(X)
print("pod", po.compare_svd(Uo, So, Vo, U.cpu().numpy(), S.cpu().numpy(), V.cpu().numpy()))
Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=1e-6)
Xr = pl.POD.reconstruct(Ur, Sr, Vr).cpu().numpy()
Xo = po.reconstruct(*po.truncate(Uo, So, Vo, r=1e-6))
rep("pod reconstruct", np.abs(Xr - Xo).max(), 1e-12)
print("launches", pl._lib.lib().pl_launch_count())
