"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dist_check.py
Compares the NCCL-composed tsqr_svd / POD.run with the CPU oracle run on the same row shards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import pyloworder_b200 as pl
import pod_oracle as po, synth

rank, size = pl.utils.init_distributed("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
ok = True
# the last case has >= 1 M rows per rank: the explicit Q is formed while exchange + small SVD run on a side stream
for (m, n, remove_mean) in ((5000, 24, False), (40000, 151, True), (3001, 64, True), (size * 70, 64, False),
                            (size * 1_000_000, 32, False)):
    X = synth.snapshots(m, n, 2021)
    shards = [X[slice(*po.worksplit(0, m, r, size))] for r in range(size)]
    r0, r1 = pl.utils.worksplit(0, m, rank, size)
    Xd = torch.from_numpy(X[r0:r1].copy()).to(dev)
    if remove_mean:
        U, S, V = pl.POD.run(Xd, remove_mean=True)
        Uo, So, Vo = po.pod_run(shards, remove_mean=True)
    else:
        U, S, V = pl.math.tsqr_svd(Xd)
        Uo, So, Vo = po.tsqr_svd(shards)
    Uh, Sh, Vh = U.cpu().numpy(), S.cpu().numpy(), V.cpu().numpy()
    sig = np.abs(Sh - So).max() / So[0]
    keep = (So / So[0] >= 1e-8)
    gaps = np.minimum(np.r_[np.inf, -np.diff(So)], np.r_[-np.diff(So), np.inf]) / So[0] >= 1e-6
    sel = keep & gaps
    ip = torch.from_numpy(np.einsum("ik,ik->k", Uo[rank], Uh)).to(dev)
    dist.all_reduce(ip)
    ipn = np.abs(ip.cpu().numpy())
    vip = np.abs(np.einsum("ki,ki->k", Vo, Vh))
    G = U.T @ U
    dist.all_reduce(G)
    orth = float((G - torch.eye(n, dtype=torch.float64, device=dev)).abs().max())
    Sall = [torch.zeros_like(S) for _ in range(size)]
    dist.all_gather(Sall, S)
    same = all(torch.equal(Sall[0], s) for s in Sall)
    Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=1e-6)
    Xr = pl.POD.reconstruct(Ur, Sr, Vr)
    Y = Xd - Xd.mean(1, keepdim=True) if remove_mean else Xd
    rm = pl.math.RMSE(Y, Xr)
    Ul = np.vstack(Uo); Xo = po.reconstruct(*po.truncate(Ul, So, Vo, r=1e-6))
    Yo = np.vstack([s - s.mean(1, keepdims=True) for s in shards]) if remove_mean else X
    rmo = po.RMSE(Yo, Xo)
    good = sig <= 1e-10 and ipn[sel].min() >= 1 - 1e-8 and vip[sel].min() >= 1 - 1e-8 and orth <= 1e-12 and same and abs(rm - rmo) <= 1e-10
    ok &= bool(good)
    if rank == 0:
        print(f"P={size} {m}x{n} mean={remove_mean}: sigma_rel={sig:.2e} mode_min={ipn[sel].min():.12f} vmode_min={vip[sel].min():.12f} "
              f"orth={orth:.2e} S_identical={same} rmse={rm:.3e}/{rmo:.3e} -> {'OK' if good else 'FAIL'}", flush=True)
# ---- randomized path: matmulp (transposed-tall GEMM + all-reduce), tsqr of the sketch, rectangular svd
for (m, n, r, q) in ((20000, 96, 12, 2), (size * 300, 40, 8, 1)):
    X = synth.snapshots(m, n, 2022)
    shards = [X[slice(*po.worksplit(0, m, k, size))] for k in range(size)]
    r0, r1 = pl.utils.worksplit(0, m, rank, size)
    Xd = torch.from_numpy(X[r0:r1].copy()).to(dev)
    U, S, V = pl.POD.run(Xd, remove_mean=True, randomized=True, r=r, q=q, seed=11)
    Uo, So, Vo = po.pod_run(shards, remove_mean=True, randomized=True, r=r, q=q, seed=11)
    Sh, Vh = S.cpu().numpy(), V.cpu().numpy()
    sig = np.abs(Sh - So).max() / So[0]
    ip = torch.from_numpy(np.einsum("ik,ik->k", Uo[rank], U.cpu().numpy())).to(dev)
    dist.all_reduce(ip)
    ipn = np.abs(ip.cpu().numpy())
    vip = np.abs(np.einsum("ki,ki->k", Vo, Vh))
    G = U.T @ U
    dist.all_reduce(G)
    orth = float((G - torch.eye(r, dtype=torch.float64, device=dev)).abs().max())
    Sall = [torch.zeros_like(S) for _ in range(size)]
    dist.all_gather(Sall, S)
    same = all(torch.equal(Sall[0], s) for s in Sall)
    good = sig <= 1e-10 and ipn.min() >= 1 - 1e-8 and vip.min() >= 1 - 1e-8 and orth <= 1e-12 and same
    ok &= bool(good)
    if rank == 0:
        print(f"P={size} randomized {m}x{n} r={r} q={q}: sigma_rel={sig:.2e} mode_min={ipn.min():.12f} vmode_min={vip.min():.12f} "
              f"orth={orth:.2e} S_identical={same} -> {'OK' if good else 'FAIL'}", flush=True)
# ---- DMD on the POD basis: projection U^T Y2 through matmulp (all-reduce), modes as interleaved-complex GEMM
for (m, n, r) in ((6000, 48, 8),):
    X = synth.dmd_waves(m, n, 3)
    shards = [X[slice(*po.worksplit(0, m, k, size))] for k in range(size)]
    r0, r1 = pl.utils.worksplit(0, m, rank, size)
    muR, muI, Phi, b = pl.DMD.run(torch.from_numpy(X[r0:r1].copy()).to(dev), r, remove_mean=True)
    muRo, muIo, Phio, bo = po.dmd_run(shards, r)
    dmu = max(np.abs(muR.cpu().numpy() - muRo).max(), np.abs(muI.cpu().numpy() - muIo).max())
    pb, pbo = Phi.cpu().numpy() * b.cpu().numpy(), Phio[rank] * bo
    dpb = torch.tensor([np.abs(pb - pbo).max() / np.abs(bo).max()], device=dev)
    dist.all_reduce(dpb, op=dist.ReduceOp.MAX)
    good = dmu <= 1e-10 and float(dpb) <= 1e-6
    ok &= bool(good)
    if rank == 0:
        print(f"P={size} DMD {m}x{n} r={r}: mu_abs={dmu:.2e} mode_amp_rel={float(dpb):.2e} -> {'OK' if good else 'FAIL'}", flush=True)
# ---- host-pointer phase calls: factor (chunked H2D || QR) -> NCCL all-gather of R -> stack SVD -> apply (GEMM || D2H)
from pyloworder_b200 import _lib
L = _lib.lib()
os.environ["PL_HOST_CHUNKS"] = "3"
for (m, n) in ((30000, 64), (9000, 33)):
    X = synth.snapshots(m, n, 2024)
    shards = [X[slice(*po.worksplit(0, m, k, size))] for k in range(size)]
    Xi = np.ascontiguousarray(shards[rank]); mi = Xi.shape[0]
    Rl = torch.empty((n, n), dtype=torch.float64, device=dev)
    Rst = torch.empty((size * n, n), dtype=torch.float64, device=dev); Wst = torch.empty_like(Rst)
    Sd = torch.empty(n, dtype=torch.float64, device=dev); Vd = torch.empty((n, n), dtype=torch.float64, device=dev)
    Uh = np.zeros((mi, n))
    rc = L.pl_tsqr_host_factor_f64(Rl.data_ptr(), Xi.ctypes.data, mi, n)
    dist.all_gather_into_tensor(Rst, Rl)
    torch.cuda.synchronize()
    rc |= L.pl_tsqr_host_stack_f64(Wst.data_ptr(), Sd.data_ptr(), Vd.data_ptr(), Rst.data_ptr(), size, n)
    rc |= L.pl_tsqr_host_apply_f64(Uh.ctypes.data, Wst[rank * n:(rank + 1) * n].data_ptr(), mi, n)
    Uo, So, Vo = po.tsqr_svd(shards)
    Sh, Vh = Sd.cpu().numpy(), Vd.cpu().numpy()
    sig = np.abs(Sh - So).max() / So[0]
    keep = (So / So[0] >= 1e-8)
    gaps = np.minimum(np.r_[np.inf, -np.diff(So)], np.r_[-np.diff(So), np.inf]) / So[0] >= 1e-6
    sel = keep & gaps
    ip = torch.from_numpy(np.einsum("ik,ik->k", Uo[rank], Uh)).to(dev)
    dist.all_reduce(ip)
    ipn = np.abs(ip.cpu().numpy())
    G = torch.from_numpy(Uh.T @ Uh).to(dev)
    dist.all_reduce(G)
    orth = float((G - torch.eye(n, dtype=torch.float64, device=dev)).abs().max())
    good = rc == 0 and sig <= 1e-10 and ipn[sel].min() >= 1 - 1e-8 and orth <= 1e-12
    ok &= bool(good)
    if rank == 0:
        print(f"P={size} host phase calls {m}x{n}: rc={rc} sigma_rel={sig:.2e} mode_min={ipn[sel].min():.12f} orth={orth:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
# ---- complex128 tsqr_svd (real embedding of every rank's shard; SPOD's call) against LAPACK on the whole matrix
for (m, n) in ((8000, 12),):
    rng = np.random.default_rng(9)
    A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) * 10.0 ** (-2.0 * np.arange(n) / n)
    r0, r1 = pl.utils.worksplit(0, m, rank, size)
    U, S, VH = pl.math.tsqr_svd(torch.from_numpy(A[r0:r1].copy()).to(dev))
    Uo, So, VHo = np.linalg.svd(A, full_matrices=False)
    sig = np.abs(S.cpu().numpy() - So).max() / So[0]
    ip = torch.view_as_real(torch.from_numpy(np.einsum("ik,ik->k", Uo[r0:r1].conj(), U.cpu().numpy())).to(dev).contiguous())
    dist.all_reduce(ip)
    ipn = np.abs(torch.view_as_complex(ip).cpu().numpy())
    rec = torch.tensor([np.abs((U.cpu().numpy() * S.cpu().numpy()) @ VH.cpu().numpy() - A[r0:r1]).max()], device=dev)
    dist.all_reduce(rec, op=dist.ReduceOp.MAX)
    good = sig <= 1e-12 and ipn.min() >= 1 - 1e-8 and float(rec) <= 1e-11 * np.abs(A).max()
    ok &= bool(good)
    if rank == 0:
        print(f"P={size} complex tsqr_svd {m}x{n}: sigma_rel={sig:.2e} mode_min={ipn.min():.12f} recon={float(rec):.2e} -> {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
