#!/usr/bin/env python
"""Benchmark of the POD / TSQR-SVD hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one `tsqr_svd` of the synthetic fp64 snapshot matrix.  At N = 1 the workload is
BASELINE.json configs[1] (8,000,000 x 512 on one B200); with N > 1 (torchrun, one rank per GPU)
every rank holds 8,000,000 rows of a global (N * 8e6) x 512 matrix (weak scaling, rows sharded
with `worksplit`) and the step contains the single NCCL all-gather of the R factors.

value          = algorithmic GFLOP/s of the whole job, F_alg = 4 m n^2, inputs
                 resident in HBM, CUDA-event timed, max over ranks.
e2e            = same metric through the host-pointer C ABI call (pl_tsqr_svd_host_f64: the
                 drop-in for the reference's dtsqr_svd), host<->device copies inside the timed region
                 (pinned host buffers; the library caches its device buffers after the first call).
roofline       = dominant kernel (caqr_update_kernel, FP64 DMMA block-reflector application),
                 per-launch CUDA-event timing from the library's profiling hooks.
cpu_baseline   = the reference's own C sources (oracle/_ref, LAPACKE+CBLAS on scipy-openblas) on the
                 host cores, on a bounded row sample of the same matrix.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_COLS = 512
ROWS_PER_GPU = 8_000_000
SEED = 2022
CPU_SAMPLE_ROWS = 150_000
PEAK_FP64_TFLOPS = 35.46          # cuBLAS DGEMM 8192^3 sustained, measured on this pool (profiles/r01_dgemm_peak.json)


def f_alg(m, n):
    return 4.0 * m * n * n


def workload_config(rows, n, size):
    """The `config` object, identical for both arms."""
    cfg_name = "BASELINE configs[1]" if (rows == ROWS_PER_GPU and n == N_COLS) else \
               ("BASELINE configs[4] 'billionaire' when run on 8 GPUs" if (rows == 125_000_000 and n == 64) else "custom shape")
    return {"workload": f"tsqr_svd of synthetic {rows}x{n} fp64 per GPU ({cfg_name}; global {rows * size}x{n}, rows sharded)",
            "rows_per_gpu": rows, "cols": n, "seed": SEED, "l2": f"inputs ({rows * n * 8 / 1e9:.1f} GB/GPU) larger than L2",
            "f_alg": "4*m*n^2"}


# ----------------------------------------------------------------------------------------------
# synthetic data (same formulas as oracle/synth.py, evaluated on the device in row chunks)
# ----------------------------------------------------------------------------------------------
def device_snapshots(torch, m_global, n, seed, r0, r1, device, out=None, chunk=250_000):
    import math
    K = min(n, 32)
    X = out if out is not None else torch.empty((r1 - r0, n), dtype=torch.float64, device=device)
    j = torch.arange(n, dtype=torch.float64, device=device)
    t = j / n
    psi = torch.stack([torch.cos(2 * math.pi * ((k + 2) // 2) * t) if k % 2 == 0 else torch.sin(2 * math.pi * ((k + 2) // 2) * t)
                       for k in range(K)])                                           # K x n
    a = torch.tensor([10.0 ** (-6.0 * k / K) for k in range(K)], dtype=torch.float64, device=device)
    M1, M2, M3 = -7046029254386353131, -4658895280553007687, -7723592293110705685       # splitmix64 constants as int64
    jj = torch.arange(n, dtype=torch.int64, device=device) * M3
    for c0 in range(r0, r1, chunk):
        c1 = min(c0 + chunk, r1)
        i = torch.arange(c0, c1, dtype=torch.int64, device=device)
        x = (i.double() + 0.5) / m_global
        kk = torch.arange(1, K + 1, dtype=torch.float64, device=device)
        phi = torch.sin(2 * math.pi * x[:, None] * kk[None, :] + 0.37 * (kk[None, :] - 1)) * a[None, :]   # rows x K
        blk = X[c0 - r0:c1 - r0]
        torch.matmul(phi, psi, out=blk)
        blk += (1.0 + 0.3 * torch.sin(2 * math.pi * x))[:, None]
        sm = (seed * M1) & ((1 << 64) - 1)
        sm = sm - (1 << 64) if sm >= (1 << 63) else sm
        z = sm ^ (i * M2)[:, None] ^ jj[None, :]
        z = z + M1
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * M2
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * M3
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        u = ((z >> 11) & ((1 << 53) - 1)).double() * (1.0 / 9007199254740992.0) - 0.5
        blk += 1e-8 * u
        del z, u, phi
    return X


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([s.strip() for s in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max([int(s[1]) for s in self.samples if s[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples for k in range(4) if len(s) > 2 + k and s[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU reference arm (the reference's own C sources, single rank, all host cores through OpenBLAS)
# ----------------------------------------------------------------------------------------------
def cpu_reference(m, n, steps, warmup):
    """Time dtsqr_svd from oracle/_ref/libpylom_ref.so on rows [0, m) of the synthetic matrix."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import synth
    path = os.path.join(ROOT, "oracle", "_ref", "libpylom_ref.so")
    kind = "reference"
    A = synth.snapshots(ROWS_PER_GPU, n, SEED, 0, m)
    if os.path.exists(path):
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        U = np.zeros((m, n)); S = np.zeros(n); V = np.zeros((n, n))

        def step():
            info = lib.dtsqr_svd(U.ctypes.data_as(dp), S.ctypes.data_as(dp), V.ctypes.data_as(dp), A.ctypes.data_as(dp),
                                 ctypes.c_int(m), ctypes.c_int(n))
            assert info == 0
    else:   # the numpy restatement of the same algorithm
        import pod_oracle as po
        kind = "port"

        def step():
            po.tsqr_svd(A)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": f_alg(m, n) / dt * 1e-9, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"first {m} rows x {n} of the synthetic matrix (seed {SEED}), dtsqr_svd single rank, "
                      f"scipy-openblas threads = all cores, {dt:.2f} s/step; rows*snapshots/s = {m * n / dt:.3e}",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, n = args.cpu_rows, args.cols
    cb = cpu_reference(m, n, max(1, args.steps), min(args.warmup, 1))
    line = {"impl": "reference", "metric": "tsqr_svd_gflops", "value": cb["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.rows, n, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import pyloworder_b200 as pl
    from pyloworder_b200 import _lib, _dev
    from pyloworder_b200.utils import parall

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank, size = parall.init_distributed("nccl") if world > 1 else (0, 1)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    n = args.cols
    m_global = args.rows * size
    r0, r1 = parall.worksplit(0, m_global, rank, size)
    m = r1 - r0

    A = device_snapshots(torch, m_global, n, SEED, r0, r1, dev)
    torch.cuda.synchronize()

    def step():
        return pl.math.tsqr_svd(A)

    def barrier():
        torch.cuda.synchronize()
        if size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    U = S = V = None
    for _ in range(args.warmup):
        U = S = V = None          # release the previous result first: U is as large as the input
        U, S, V = step()
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    l0 = L.pl_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        U = S = V = None
        U, S, V = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (L.pl_launch_count() - l0) // args.steps
    if size > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None

    # ---- full-size sanity (size-independent properties; no CPU oracle at this scale)
    chk = {}
    sub = slice(0, min(m, 200_000))
    G = U[sub].T @ U[sub]
    chk["s_desc"] = bool((S[:-1] >= S[1:]).all().item())
    chk["VVt_minus_I_max"] = float((V @ V.T - torch.eye(n, dtype=torch.float64, device=dev)).abs().max().item())
    rec = (U[sub] * S) @ V
    chk["recon_rel_sample"] = float(((rec - A[sub]).norm() / A[sub].norm()).item())
    UtU = U.T @ U
    if size > 1:
        dist.all_reduce(UtU)
    chk["UtU_minus_I_max"] = float((UtU - torch.eye(n, dtype=torch.float64, device=dev)).abs().max().item())
    del G, rec, UtU

    U = None
    torch.cuda.empty_cache()
    # ---- per-kernel-class timing of one extra step (profiling hooks; not part of the timed steps)
    L.pl_profile_enable(1)
    Up, _, _ = step()
    torch.cuda.synchronize()
    del Up
    L.pl_profile_enable(0)
    NC = 7
    msb = (ctypes.c_double * NC)(); cnt = (ctypes.c_int64 * NC)()
    L.pl_profile_read(ctypes.cast(msb, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), NC)
    names = ["copy_center", "panel", "update_factor", "update_formq", "gemm", "svd_small", "misc"]
    phases = {names[i]: {"ms": round(msb[i], 3), "launches": int(cnt[i])} for i in range(NC)}
    # dominant kernel: caqr_update_kernel.  Algorithmic flops of one block-reflector application
    # = 4 * rows * NB * cols (W = V^T C and C -= V W').  Factor pass: cols = trailing columns of each panel;
    # form-Q pass: trailing + own panel columns.
    npad = -(-n // 32) * 32
    Kp = npad // 32
    fl_f = sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * (p + 1)) for p in range(Kp))
    fl_q = sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * p) for p in range(Kp))
    upd_ms = msb[2] + msb[3]
    upd_launch = int(cnt[2] + cnt[3])
    ach = (fl_f + fl_q) / (upd_ms * 1e-3) * 1e-12 if upd_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "caqr_update_kernel (FP64 DMMA)", "achieved": ach, "peak": PEAK_FP64_TFLOPS,
                "unit": "TFLOP/s", "frac": ach / PEAK_FP64_TFLOPS,
                "peak_source": "measured cuBLAS DGEMM 8192^3 on this pool (profiles/r01_dgemm_peak.json); MEASURED_PEAKS.json has no FP64 entry",
                "avg_launch_ms": upd_ms / max(upd_launch, 1), "launches_per_step": upd_launch,
                "share_of_step": upd_ms / sum(msb), "traffic": None,
                "timing": "CUDA events around every launch of one extra step",
                "ncu_traffic": {"launch": "first factor-pass launch of the 1,000,000 x 512 probe (15 chunks x 977 strips), profiles/r01_ncu_update.txt",
                                "dram_bytes": 7.948e9, "algorithmic_bytes": 7.94e9}}

    # ---- e2e through the host-pointer C ABI (rank 0 of N; every rank does its own shard)
    e2e = None
    S_dev = S.cpu() if size == 1 else None
    del S, V
    # host memory: 2 pinned buffers of m_e2e x n per rank; with several ranks on one node the e2e sample is capped
    e2e_rows = args.e2e_rows if args.e2e_rows > 0 else (ROWS_PER_GPU if size == 1 else (2_000_000 if size <= 4 else 1_000_000))
    m_e2e = min(m, e2e_rows)
    alloc_err = None
    try:
        host_in = torch.empty((m_e2e, n), dtype=torch.float64, pin_memory=True)
        host_in.copy_(A[:m_e2e])
        host_U = torch.empty((m_e2e, n), dtype=torch.float64, pin_memory=True)
        host_S = torch.empty(n, dtype=torch.float64); host_V = torch.empty((n, n), dtype=torch.float64)
    except Exception as ex:  # pinned host memory too small
        alloc_err = str(ex)[:200]
    del A
    _dev.free_workspaces()
    torch.cuda.empty_cache()
    if size > 1:   # the e2e leg is collective: every rank runs it or none does
        flag = torch.tensor([1.0 if alloc_err else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag.item()) > 0 and not alloc_err:
            alloc_err = "pinned host allocation failed on another rank"
    try:
        if alloc_err:
            raise RuntimeError(alloc_err)
        if size == 1:
            def e2e_step():
                rc = L.pl_tsqr_svd_host_f64(host_U.data_ptr(), host_S.data_ptr(), host_V.data_ptr(), host_in.data_ptr(), m_e2e, n)
                _lib.check(rc, "pl_tsqr_svd_host_f64")
        else:
            # P ranks from host memory: local chunked pipeline -> NCCL all-gather of the n x n R's -> SVD of the stack
            # (redundant on every rank) -> local chunked back-multiply with this rank's block of Q2 Ur
            Rl = torch.empty((n, n), dtype=torch.float64, device=dev)
            Rst = torch.empty((size * n, n), dtype=torch.float64, device=dev)
            Wst = torch.empty((size * n, n), dtype=torch.float64, device=dev)
            Sd = torch.empty(n, dtype=torch.float64, device=dev); Vd = torch.empty((n, n), dtype=torch.float64, device=dev)
            _, wp2, wb2 = _dev.workspace(L.pl_qr_workspace_bytes(size * n, n), "stack", dev)
            def e2e_step():
                _lib.check(L.pl_tsqr_host_factor_f64(Rl.data_ptr(), host_in.data_ptr(), m_e2e, n), "pl_tsqr_host_factor_f64")
                dist.all_gather_into_tensor(Rst, Rl)
                _lib.check(L.pl_tsqr_svd_f64(Wst.data_ptr(), Sd.data_ptr(), Vd.data_ptr(), Rst.data_ptr(), size * n, n, wp2, wb2,
                                             _dev.stream()), "pl_tsqr_svd_f64")
                host_S.copy_(Sd); host_V.copy_(Vd)           # synchronises the torch stream: W is ready
                _lib.check(L.pl_tsqr_host_apply_f64(host_U.data_ptr(), Wst[rank * n:(rank + 1) * n].data_ptr(), m_e2e, n),
                           "pl_tsqr_host_apply_f64")
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if size > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": f_alg(m_e2e * size, n) / dt * 1e-9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": m_e2e * n * 8, "d2h_bytes_per_step": m_e2e * n * 8 + n * 8 + n * n * 8,
               "rows_per_gpu": m_e2e, "ms_per_step": dt * 1e3,
               "path": "pl_tsqr_svd_host_f64 (host pointers; row-chunk pipeline H2D || factor+Q, GEMM || D2H inside the timed region, device buffers cached by the library after the warm-up call)" if size == 1
                       else "pl_tsqr_host_factor_f64 -> NCCL all-gather of R -> pl_tsqr_svd_f64 on the stack -> pl_tsqr_host_apply_f64 (host pointers; chunked H2D / D2H inside the timed region)"}
        if size == 1 and m_e2e == m:
            e2e["s_rel_diff_vs_device_path"] = float((host_S - S_dev).abs().max() / S_dev[0])
    except Exception as ex:  # host memory too small etc.
        e2e = {"value": None, "unit": "GFLOP/s", "error": str(ex)[:200]}

    if rank == 0:
        cb = None
        if size == 1 and not args.no_cpu:
            cb = cpu_reference(args.cpu_rows, n, 1, 0)
        value = f_alg(m_global, n) / (ms * 1e-3) * 1e-9
        line = {"metric": "tsqr_svd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": workload_config(args.rows, n, size),
                "rows_snapshots_per_s": m_global * n / (ms * 1e-3),
                "frac_of_fp64_roofline": value * 1e-3 / (PEAK_FP64_TFLOPS * size),
                "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline, "phases_ms": phases,
                "checks": chk}
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU")
    ap.add_argument("--cols", type=int, default=N_COLS)
    ap.add_argument("--cpu-rows", type=int, default=CPU_SAMPLE_ROWS)
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per GPU of the e2e leg (0: full shard on 1 GPU, 2M per GPU on 2-4 GPUs, 1M on 8: pinned host memory)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
