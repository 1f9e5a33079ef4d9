"""P-rank CPU run of the reference's TSQR-SVD  --  TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).

The reference runs `tsqr_svd` as P MPI ranks (`mpirun -np p`, Examples/scalability/generate_MN5_GPU.sh:53); this
image has no MPI, so the ranks here are P operating-system processes and a send/recv pair of the butterfly
(pyLOM/vmmath/src/svd.c:602-669, pyLOM/vmmath/svd.py:67-115) is a pipe between two of them.  Every rank executes
the per-rank control flow of `dtsqr_svd` (svd.c:678-712): local `dqr`, reduction levels with partner
`rank ^ level`, broadcast levels composing QW, `Qi = Q1i QW`, `dsvd(R)` on every rank, `Ui = Qi Ur`.
The local arithmetic is the reference's own compiled C (`dqr`, `dsvd`, `dmatmul` of oracle/_ref/libpylom_ref.so:
LAPACKE + CBLAS on scipy-openblas) when that library was built, else numpy (kind "port").
BLAS threads per rank = cores // P, set through the environment BEFORE the BLAS is loaded (torchrun exports
OMP_NUM_THREADS=1, which would otherwise throttle the arm).

Used by `bench.py --impl reference` / the `cpu_baseline` leg and by tests/test_oracle.py (P-rank result == the
level-synchronous simulation in pod_oracle.tsqr_svd).
"""
import ctypes
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpylom_ref.so")


def _next_pow2(p):
    q = 1
    while q < p:
        q <<= 1
    return q


class _Kernels:
    """Local dense kernels of one rank: the compiled reference when available, numpy otherwise."""

    def __init__(self, use_ref):
        import numpy as np
        self.np = np
        self.lib = None
        if use_ref and os.path.exists(REF_SO):
            self.lib = ctypes.CDLL(REF_SO)
            self.dp = ctypes.POINTER(ctypes.c_double)
        self.kind = "reference" if self.lib is not None else "port"

    def _p(self, a):
        return a.ctypes.data_as(self.dp)

    def qr(self, A):
        np = self.np
        m, n = A.shape
        if self.lib is None:
            return np.linalg.qr(A)
        Q = np.empty((m, n)); R = np.empty((n, n))
        info = self.lib.dqr(self._p(Q), self._p(R), self._p(np.ascontiguousarray(A)), ctypes.c_int(m), ctypes.c_int(n))
        assert info == 0, info
        return Q, R

    def svd(self, R):
        np = self.np
        n = R.shape[0]
        if self.lib is None:
            return np.linalg.svd(R, full_matrices=False)
        U = np.empty((n, n)); S = np.empty(n); VT = np.empty((n, n))
        info = self.lib.dsvd(self._p(U), self._p(S), self._p(VT), self._p(np.ascontiguousarray(R.copy())), ctypes.c_int(n), ctypes.c_int(n))
        assert info == 0, info
        return U, S, VT

    def matmul(self, A, B):
        np = self.np
        if self.lib is None:
            return A @ B
        m, k = A.shape
        n = B.shape[1]
        C = np.empty((m, n))
        self.lib.dmatmul(self._p(C), self._p(np.ascontiguousarray(A)), self._p(np.ascontiguousarray(B)),
                         ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k))
        return C


def tsqr_svd_rank(Ai, rank, size, links, K):
    """One rank of the reference's tsqr_svd.  links[p] = duplex pipe to partner p."""
    np = K.np
    n = Ai.shape[1]
    Q1, R = K.qr(Ai)
    levels = _next_pow2(size).bit_length() - 1
    kept = []                                  # (level bit, stacked-QR Q factor) of every reduction this rank performed
    bit = 1
    for _ in range(levels):                    # ---- reduction towards rank 0
        partner = rank ^ bit
        if rank & bit:
            if partner < size:
                links[partner].send(R)
        elif partner < size:
            Rb = links[partner].recv()
            Q2, R = K.qr(np.vstack((R, Rb)))
            kept.append((bit, Q2))
        else:
            kept.append((bit, None))
        bit <<= 1
    QW = np.eye(n)
    bit = 1 << (levels - 1) if levels else 0
    mask = bit - 1 if levels else 0
    for _ in range(levels):                    # ---- broadcast of R and of the composed Q correction
        if rank & mask == 0:
            partner = rank ^ bit
            if rank & bit:
                if partner < size:
                    msg = links[partner].recv()
                    R, QW = msg[:n], msg[n:]
            else:
                q2 = dict(kept).get(bit)
                if q2 is not None:
                    Q2W = K.matmul(q2, QW)
                    links[partner].send(np.vstack((R, Q2W[n:])))
                    QW = Q2W[:n]
        bit >>= 1
        mask >>= 1
    Qi = K.matmul(Q1, QW)
    Ur, S, VT = K.svd(R)
    return K.matmul(Qi, Ur), S, VT


def _worker(rank, size, links, m_total, n, seed, m_global, steps, warmup, use_ref, barrier, out, keep):
    sys.path.insert(0, HERE)
    import numpy as np
    import synth
    from pod_oracle import worksplit
    K = _Kernels(use_ref)
    r0, r1 = worksplit(0, m_total, rank, size)
    Ai = synth.snapshots(m_global, n, seed, r0, r1)
    res = None
    for _ in range(warmup):
        barrier.wait()
        res = tsqr_svd_rank(Ai, rank, size, links, K)
    times = []
    for _ in range(steps):
        barrier.wait()
        t0 = time.perf_counter()
        res = tsqr_svd_rank(Ai, rank, size, links, K)
        barrier.wait()
        times.append(time.perf_counter() - t0)
    msg = {"rank": rank, "times": times, "kind": K.kind, "rows": r1 - r0}
    if keep and res is not None:
        msg["U"], msg["S"], msg["VT"] = res
    out.put(msg)


def run(P, m_total, n, seed, m_global=None, steps=1, warmup=1, threads_per_rank=None, use_ref=True, keep=False):
    """Run the P-rank tsqr_svd on the first m_total rows of the synthetic matrix (rows split with worksplit).
    Returns {"seconds": mean step time (max over ranks per step), "kind", "threads_per_rank", "cores", results...}."""
    cores = os.cpu_count() or 1
    tpr = threads_per_rank or max(1, cores // P)
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = str(tpr)               # inherited by the spawned ranks, read when their BLAS loads
    try:
        ctx = mp.get_context("spawn")
        pipes = {}
        for a in range(P):
            for b in range(a + 1, P):
                if bin(a ^ b).count("1") == 1:  # butterfly partners only
                    pipes[(a, b)] = ctx.Pipe(duplex=True)
        barrier = ctx.Barrier(P)
        out = ctx.Queue()
        procs = []
        for r in range(P):
            links = {}
            for (a, b), (ca, cb) in pipes.items():
                if a == r:
                    links[b] = ca
                elif b == r:
                    links[a] = cb
            p = ctx.Process(target=_worker, args=(r, P, links, m_total, n, seed, m_global or m_total, steps, warmup,
                                                  use_ref, barrier, out, keep))
            p.start()
            procs.append(p)
        msgs = [out.get(timeout=3600) for _ in range(P)]
        for p in procs:
            p.join()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    msgs.sort(key=lambda d: d["rank"])
    per_step = [max(d["times"][s] for d in msgs) for s in range(steps)]
    res = {"seconds": sum(per_step) / len(per_step), "kind": msgs[0]["kind"], "threads_per_rank": tpr, "cores": cores,
           "ranks": P, "rows": [d["rows"] for d in msgs]}
    if keep:
        res["U"] = [d["U"] for d in msgs]
        res["S"] = [d["S"] for d in msgs]
        res["VT"] = [d["VT"] for d in msgs]
    return res


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    r = run(P, m, n, 2022, steps=2, warmup=1)
    print({k: v for k, v in r.items() if k not in ("U", "S", "VT")})
