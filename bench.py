#!/usr/bin/env python
"""Benchmark of the POD / TSQR-SVD hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one `tsqr_svd` of the synthetic fp64 snapshot matrix.  At N = 1 the workload is
BASELINE.json configs[1] (8,000,000 x 512 on one B200); with N > 1 (torchrun, one rank per GPU)
every rank holds 8,000,000 rows of a global (N * 8e6) x 512 matrix (weak scaling, rows sharded
with `worksplit`) and the step contains the single NCCL all-gather of the R factors.

value          = algorithmic GFLOP/s of the whole job, F_alg = 4 m n^2, inputs
                 resident in HBM, CUDA-event timed, max over ranks.
parity         = BEFORE the timed loop: down-scaled twins (N*20,000 x 512 tsqr_svd, N*40,000 x 64
                 POD.run with centering) compared with the CPU oracle (oracle/pod_oracle.py, the
                 reference's P-rank butterfly) on the same worksplit shards, north-star tolerances.
e2e            = same metric through the host-pointer C ABI call (pl_tsqr_svd_host_f64: the
                 drop-in for the reference's dtsqr_svd), host<->device copies inside the timed region
                 (pinned host buffers; the library caches its device buffers after the first call);
                 `pageable` = the same call on plain numpy arrays (N = 1).
roofline       = dominant kernel of the step (caqr_update2_kernel, FP64 DMMA block-reflector application),
                 per-launch CUDA-event timing from the library's profiling hooks.
other_configs  = the other BASELINE shapes as per-GPU shards: cfg5 (125,000,000 x 64 tsqr_svd; on 8 GPUs this
                 IS config 5), cfg3 (24,000,000 x 256 POD.run(remove_mean)), cfg4 (2,000,000 x 1000 DMD),
                 cfg1 (89,351 x 151 POD, one GPU), and the strong-scaling reading of cfg2 (8 M rows / N).
cpu_baseline   = the reference's CPU path (P = N ranks as processes, the reference's compiled C kernels from
                 oracle/_ref, BLAS threads = cores // P) on a bounded row sample of the same matrix.
"""
import argparse
import ctypes
import json
import math
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


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6553.3


def f_alg(m, n):
    return 4.0 * m * n * n


def t_roof_ms(m, n, center=False):
    """Per-GPU roofline time of one tsqr_svd / POD.run: the slower of FP64 flops and HBM bytes."""
    fl = f_alg(m, n) + (2.0 * m * n if center else 0.0)
    by = 32.0 * m * n + ((16.0 * m * n + 8.0 * m) if center else 0.0)
    return max(fl / (PEAK_FP64_TFLOPS * 1e12), by / (hbm_peak_gbs() * 1e9)) * 1e3


def cpu_sample_rows(n, steps, warmup):
    """Rows of the CPU sample: ~150,000 x 512 (4.6 s on 16 cores) for short runs, scaled down so that
    (steps + warmup) samples stay within about a minute."""
    k = max(1, steps + warmup)
    rows = min(CPU_SAMPLE_ROWS, CPU_SAMPLE_ROWS * 12 // k)
    return max(rows, 16 * n)


def workload_config(rows, n, size, steps=3, warmup=1):
    """The `config` object, identical for both arms."""
    cfg_name = "BASELINE configs[1]" if (rows == ROWS_PER_GPU and n == N_COLS) else \
               ("BASELINE configs[4] 'billionaire' when run on 8 GPUs" if (rows == 125_000_000 and n == 64) else "custom shape")
    return {"workload": f"tsqr_svd of synthetic {rows}x{n} fp64 per GPU ({cfg_name}; global {rows * size}x{n}, rows sharded)",
            "rows_per_gpu": rows, "cols": n, "seed": SEED, "l2": f"inputs ({rows * n * 8 / 1e9:.1f} GB/GPU) larger than L2",
            "f_alg": "4*m*n^2",
            "reference_arm_sample": f"down-scaled twin: the first {cpu_sample_rows(n, steps, warmup)} rows of the same matrix split over "
                                    f"{size} CPU rank(s) (cost is linear in m; the reference's int32 m*n*8 arithmetic caps a rank below 2^31 bytes), "
                                    "rate-extrapolated"}


# ----------------------------------------------------------------------------------------------
# synthetic data (same formulas as oracle/synth.py, evaluated on the device in row chunks)
# ----------------------------------------------------------------------------------------------
def device_snapshots(torch, m_global, n, seed, r0, r1, device, out=None, chunk=250_000):
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
# CPU reference arm: the reference's P-rank CPU path (ranks = processes, oracle/ref_ranks.py)
# ----------------------------------------------------------------------------------------------
def cpu_reference(P, m_total, n, steps, warmup):
    """Time the reference's tsqr_svd on the first m_total rows of the synthetic matrix, split over P CPU ranks."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_ranks
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))        # the e2e leg may have narrowed this process to one NUMA node
    except Exception:
        pass
    per = max(m_total // P, 4 * n)
    m_total = per * P
    r = ref_ranks.run(P, m_total, n, SEED, m_global=ROWS_PER_GPU * P, steps=steps, warmup=warmup)
    dt = r["seconds"]
    return {"value": f_alg(m_total, n) / dt * 1e-9, "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
            "ranks": P, "threads_per_rank": r["threads_per_rank"],
            "sample": f"first {m_total} rows x {n} of the synthetic matrix (seed {SEED}) split over {P} CPU rank(s) "
                      f"({per} rows each, ranks = processes, butterfly over pipes: no MPI in the image), local kernels = the reference's "
                      f"compiled dqr/dsvd/dmatmul (LAPACKE + CBLAS, scipy-openblas 0.3.30) with {r['threads_per_rank']} BLAS threads per rank "
                      f"on {r['cores']} cores, {warmup} warm-up + {steps} timed, {dt:.2f} s/step; rows*snapshots/s = {m_total * n / dt:.3e}"
                      if r["kind"] == "reference" else
                      f"first {m_total} rows x {n} (seed {SEED}) over {P} CPU rank(s), numpy port of the reference algorithm "
                      f"(oracle/_ref not built), {r['threads_per_rank']} threads per rank on {r['cores']} cores, {dt:.2f} s/step",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = int(os.environ.get("WORLD_SIZE", "1"))
    n = args.cols
    m = args.cpu_rows if args.cpu_rows > 0 else cpu_sample_rows(n, args.steps, args.warmup)
    cb = cpu_reference(P, m, n, max(1, args.steps), args.warmup)
    line = {"impl": "reference", "metric": "tsqr_svd_gflops", "value": cb["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.rows, n, P, args.steps, args.warmup),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "ranks", "threads_per_rank")},
            "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    pass


def barrier(cx):
    cx.torch.cuda.synchronize()
    if cx.size > 1:
        cx.dist.barrier()
    cx.torch.cuda.synchronize()


def max_over_ranks(cx, v):
    if cx.size == 1:
        return v
    t = cx.torch.tensor([v], dtype=cx.torch.float64, device=cx.dev)
    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    return float(t.item())


def timed(cx, step, steps, warmup):
    """W untimed + K timed steps between barrier + synchronize; CUDA events; max over ranks.  Returns (ms, last result)."""
    torch = cx.torch
    res = None
    for _ in range(warmup):
        res = None                     # release the previous result first: U is as large as the input
        res = step()
    barrier(cx)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = None
        res = step()
    e1.record()
    barrier(cx)
    return max_over_ranks(cx, e0.elapsed_time(e1) / steps), res


def release(cx):
    from pyloworder_b200 import _dev
    _dev.free_workspaces()
    cx.torch.cuda.empty_cache()


def orth_checks(cx, U, S, V, A=None, center_mean=None):
    torch, dist = cx.torch, cx.dist
    n = V.shape[0]
    k = U.shape[1]
    chk = {"s_desc": bool((S[:-1] >= S[1:]).all().item())}
    eye = torch.eye(k, dtype=torch.float64, device=cx.dev)
    chk["VVt_minus_I_max"] = float((V @ V.T - eye[:V.shape[0], :V.shape[0]]).abs().max().item())
    UtU = torch.zeros((k, k), dtype=torch.float64, device=cx.dev)
    step = 4_000_000
    for r0 in range(0, U.shape[0], step):
        blk = U[r0:r0 + step]
        UtU += blk.T @ blk
    if cx.size > 1:
        dist.all_reduce(UtU)
    chk["UtU_minus_I_max"] = float((UtU - eye).abs().max().item())
    if A is not None:
        sub = slice(0, min(U.shape[0], 200_000))
        ref = A[sub] if center_mean is None else A[sub] - center_mean[sub, None]
        rec = (U[sub] * S) @ V
        chk["recon_rel_sample"] = float(((rec - ref).norm() / ref.norm()).item())
    return chk


# ---- parity against the CPU oracle on down-scaled twins (every N, before the timed loop) -------------
def parity_block(cx):
    import numpy as np
    torch, dist = cx.torch, cx.dist
    pl = cx.pl
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pod_oracle as po
    import synth
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    size, rank, dev = cx.size, cx.rank, cx.dev
    cases = []
    all_ok = True
    for (name, m, n, center) in (("tsqr_svd", size * 20_000, 512, False), ("POD.run(remove_mean=True)", size * 40_000, 64, True)):
        r0, r1 = pl.utils.worksplit(0, m, rank, size)
        Xd = device_snapshots(torch, m, n, 2021, r0, r1, dev)
        if center:
            U, S, V = pl.POD.run(Xd, remove_mean=True)
        else:
            U, S, V = pl.math.tsqr_svd(Xd)
        Ur, Sr, Vr = pl.POD.truncate(U, S, V, r=1e-6)
        Xr = pl.POD.reconstruct(Ur, Sr, Vr)
        Y = Xd - Xd.mean(1, keepdim=True) if center else Xd
        rm = float(pl.math.RMSE(Y, Xr))
        # oracle on rank 0 (all host cores: torchrun pins OMP_NUM_THREADS=1), results broadcast from there
        Uo_d = torch.empty((m, n), dtype=torch.float64, device=dev)
        SV_d = torch.empty((n + 1) * n + 1, dtype=torch.float64, device=dev)
        t_or = 0.0
        if rank == 0:
            t0 = time.perf_counter()
            X = synth.snapshots(m, n, 2021)
            shards = [X[slice(*po.worksplit(0, m, r, size))] for r in range(size)]
            lim = threadpool_limits(limits=os.cpu_count()) if threadpool_limits else None
            if center:
                Uo, So, Vo = po.pod_run(shards, remove_mean=True)
                Yo = np.vstack([s - s.mean(1, keepdims=True) for s in shards])
            else:
                Uo, So, Vo = po.tsqr_svd(shards)
                Yo = X
            Ul = np.vstack(Uo)
            rmo = po.RMSE(Yo, po.reconstruct(*po.truncate(Ul, So, Vo, r=1e-6)))
            if lim is not None:
                lim.restore_original_limits() if hasattr(lim, "restore_original_limits") else None
            Uo_d.copy_(torch.from_numpy(Ul))
            SV_d.copy_(torch.from_numpy(np.concatenate([So, Vo.ravel(), [rmo]])))
            t_or = time.perf_counter() - t0
            gen_diff = float((torch.from_numpy(X[r0:r1]).to(dev) - Xd).abs().max().item())   # device generator == oracle/synth.py
        if size > 1:
            dist.broadcast(Uo_d, 0)
            dist.broadcast(SV_d, 0)
        So = SV_d[:n]; Vo = SV_d[n:n + n * n].view(n, n); rmo = float(SV_d[-1].item())
        sig = float(((S - So).abs().max() / So[0]).item())
        rel = So / So[0]
        big = rel >= 1e-6
        sig_each = float((((S - So).abs() / So)[big]).max().item()) if bool(big.any()) else 0.0
        dS = So[:-1] - So[1:]
        inf = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
        gap = torch.minimum(torch.cat([inf, dS]), torch.cat([dS, inf])) / So[0]
        sel = (rel >= 1e-8) & (gap >= 1e-6)
        ip = (Uo_d[r0:r1] * U).sum(0)
        if size > 1:
            dist.all_reduce(ip)
        ip = ip.abs()
        vip = (Vo * V).sum(1).abs()
        same = True
        if size > 1:
            Sall = [torch.zeros_like(S) for _ in range(size)]
            dist.all_gather(Sall, S)
            same = all(torch.equal(Sall[0], s) for s in Sall)
        chk = orth_checks(cx, U, S, V)
        mode_min = float(ip[sel].min().item()) if bool(sel.any()) else 1.0
        vmode_min = float(vip[sel].min().item()) if bool(sel.any()) else 1.0
        ok = (sig <= 1e-10 and sig_each <= 1e-10 and mode_min >= 1 - 1e-8 and vmode_min >= 1 - 1e-8 and abs(rm - rmo) <= 1e-10
              and same and chk["UtU_minus_I_max"] <= 1e-12)
        all_ok &= bool(ok)
        c = {"case": f"{name} {m}x{n} on {size} rank(s), oracle = reference butterfly on the same worksplit shards",
             "sigma_rel_to_s1": sig, "sigma_rel_each_ge_1e-6": sig_each, "modes_checked": int(sel.sum().item()),
             "mode_min_abs_inner": mode_min, "vmode_min_abs_inner": vmode_min, "rmse": rm, "rmse_oracle": rmo,
             "rmse_abs_diff": abs(rm - rmo), "S_bit_identical_across_ranks": bool(same),
             "UtU_minus_I_max": chk["UtU_minus_I_max"], "ok": bool(ok)}
        if rank == 0:
            c["oracle_seconds"] = round(t_or, 2)
            c["input_max_abs_diff_device_vs_oracle_generator"] = gen_diff
        cases.append(c)
        del Xd, U, S, V, Ur, Sr, Vr, Xr, Y, Uo_d, SV_d
        release(cx)
    return {"tolerances": {"sigma": 1e-10, "mode": "1-1e-8", "rmse_abs": 1e-10}, "ok": bool(all_ok), "cases": cases}


# ---- other BASELINE configs as per-GPU shards ---------------------------------------------------------
def other_config(cx, name, rows, n, kind, steps=5, warmup=3, m_global=None, seed=2025):
    """Time one per-GPU shard; returns the dict that goes under other_configs[name]."""
    torch, pl = cx.torch, cx.pl
    size, rank, dev = cx.size, cx.rank, cx.dev
    m_global = m_global or rows * size
    r0, r1 = pl.utils.worksplit(0, m_global, rank, size)
    m = r1 - r0
    out = {"rows_per_gpu": m, "cols": n, "global_rows": m_global, "op": kind}
    try:
        A = device_snapshots(torch, m_global, n, seed, r0, r1, dev)
        torch.cuda.synchronize()
        L = cx.L
        l0 = L.pl_launch_count()
        if kind == "tsqr_svd":
            step = lambda: pl.math.tsqr_svd(A)
        elif kind == "POD.run(remove_mean=True)":
            step = lambda: pl.POD.run(A, remove_mean=True)
        elif kind == "DMD.run":
            step = lambda: pl.DMD.run(A, 1e-6, remove_mean=False)
        else:
            raise ValueError(kind)
        ms, res = timed(cx, step, steps, warmup)
        out["gpu_launches"] = int((L.pl_launch_count() - l0) // (steps + warmup))
        nn = n - 1 if kind == "DMD.run" else n
        center = kind.startswith("POD")
        fl = f_alg(m_global, nn) + (2.0 * m_global * n if center else 0.0)
        troof = t_roof_ms(m, nn, center)
        out.update({"ms": ms, "steps": steps, "warmup": warmup, "gflops_alg": fl / (ms * 1e-3) * 1e-9,
                    "rows_snapshots_per_s": m_global * n / (ms * 1e-3), "t_roof_ms": troof, "frac_of_roofline": troof / ms,
                    "roofline_bound": "fp64" if f_alg(m, nn) / (PEAK_FP64_TFLOPS * 1e12) >= 32.0 * m * nn / (hbm_peak_gbs() * 1e9) else "hbm"})
        if kind != "DMD.run":
            U, S, V = res
            mean = A.mean(1) if center else None
            out["checks"] = orth_checks(cx, U, S, V, A, mean)
            del U, S, V
        else:
            out["checks"] = {"n_modes": int(res[0].numel())}
        del res, A
    except Exception as ex:
        out["error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
    release(cx)
    return out


def numa_bind(torch, local):
    """Pin this rank's host threads to the cores next to its GPU, so that pinned staging buffers are first-touched on
    the GPU's NUMA node (8 ranks on one host otherwise pile their buffers on one socket)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


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
    cx = Ctx()
    cx.torch, cx.dist, cx.pl, cx.L, cx.size, cx.rank, cx.dev = torch, dist, pl, L, size, rank, dev
    n = args.cols
    m_global = args.rows * size
    r0, r1 = parall.worksplit(0, m_global, rank, size)
    m = r1 - r0
    skip = set(s for s in args.skip.split(",") if s)

    # ---- parity twins against the oracle (before anything is timed)
    parity = None
    if "parity" not in skip:
        try:
            parity = parity_block(cx)
        except Exception as ex:
            parity = {"ok": False, "error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    A = device_snapshots(torch, m_global, n, SEED, r0, r1, dev)
    torch.cuda.synchronize()

    def step():
        return pl.math.tsqr_svd(A)

    clocks = Clocks(local)
    U = S = V = None
    for _ in range(args.warmup):
        U = S = V = None
        U, S, V = step()
    barrier(cx)
    if rank == 0:
        clocks.start()
    l0 = L.pl_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        U = S = V = None
        U, S, V = step()
    e1.record()
    barrier(cx)
    ms = max_over_ranks(cx, e0.elapsed_time(e1) / args.steps)
    launches = (L.pl_launch_count() - l0) // args.steps
    clk = clocks.stop() if rank == 0 else None

    # ---- full-size sanity (size-independent properties; no CPU oracle at this scale)
    chk = orth_checks(cx, U, S, V, A)
    U = None
    torch.cuda.empty_cache()
    # ---- per-kernel-class timing of one extra step (profiling hooks; not part of the timed steps)
    L.pl_profile_enable(1)
    Up, _, _ = step()
    torch.cuda.synchronize()
    del Up
    L.pl_profile_enable(0)
    NC = 8
    msb = (ctypes.c_double * NC)(); cnt = (ctypes.c_int64 * NC)()
    L.pl_profile_read(ctypes.cast(msb, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), NC)
    names = ["copy_center", "panel", "update_factor", "update_formq", "gemm", "svd_small", "misc", "tsqr_small"]
    phases = {names[i]: {"ms": round(msb[i], 3), "launches": int(cnt[i])} for i in range(NC)}
    # dominant kernel: caqr_update2_kernel.  Algorithmic flops of one block-reflector application
    # = 4 * rows * NB * cols (W = V^T C and C -= V W').  Factor pass: cols = trailing columns of each panel;
    # form-Q pass: trailing + own panel columns.
    npad = -(-n // 32) * 32
    Kp = npad // 32
    fl_f = sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * (p + 1)) for p in range(Kp))
    fl_q = sum(4.0 * (m - 32 * p) * 32 * (npad - 32 * p) for p in range(Kp))
    upd_ms = msb[2] + msb[3]
    upd_launch = int(cnt[2] + cnt[3])
    ach = (fl_f + fl_q) / (upd_ms * 1e-3) * 1e-12 if upd_ms > 0 else 0.0
    traffic = None
    traffic_note = None
    try:   # per-launch DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this command
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_update_traffic.json")))
        traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "caqr_update2_kernel (FP64 DMMA, m8n8k4)", "achieved": ach, "peak": PEAK_FP64_TFLOPS,
                "unit": "TFLOP/s", "frac": ach / PEAK_FP64_TFLOPS,
                "peak_source": "measured cuBLAS DGEMM 8192^3 on this pool (profiles/r01_dgemm_peak.json); MEASURED_PEAKS.json has no FP64 entry",
                "avg_launch_ms": upd_ms / max(upd_launch, 1), "launches_per_step": upd_launch,
                "share_of_step": upd_ms / max(sum(msb), 1e-9), "traffic": traffic, "traffic_source": traffic_note,
                "timing": "CUDA events around every launch of one extra step"}

    # ---- e2e through the host-pointer C ABI (every rank does its own shard)
    e2e = None
    S_dev = S.cpu() if size == 1 else None
    del S, V
    e2e_rows = args.e2e_rows if args.e2e_rows > 0 else ROWS_PER_GPU
    m_e2e = min(m, e2e_rows)
    if args.e2e_rows <= 0:     # full shard only when the host has the memory for 2 pinned buffers per rank
        try:
            import psutil
            avail = psutil.virtual_memory().available
            while m_e2e > 250_000 and 2.0 * m_e2e * n * 8 * size > 0.6 * avail:
                m_e2e //= 2
        except Exception:
            if size > 1:
                m_e2e = min(m_e2e, 2_000_000 if size <= 4 else 1_000_000)
    numa = numa_bind(torch, local) if "numa" not in skip else None
    alloc_err = None
    try:
        if "e2e" in skip:
            raise RuntimeError("skipped (--skip e2e)")
        host_in = torch.empty((m_e2e, n), dtype=torch.float64, pin_memory=True)
        host_in.copy_(A[:m_e2e])
        host_U = torch.empty((m_e2e, n), dtype=torch.float64, pin_memory=True)
        host_S = torch.empty(n, dtype=torch.float64); host_V = torch.empty((n, n), dtype=torch.float64)
    except Exception as ex:  # pinned host memory too small
        alloc_err = str(ex)[:200]
    del A
    release(cx)
    if size > 1:   # the e2e leg is collective: every rank runs it or none does
        flag = torch.tensor([1.0 if alloc_err else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag.item()) > 0 and not alloc_err:
            alloc_err = "pinned host allocation failed on another rank"
    try:
        if alloc_err:
            raise RuntimeError(alloc_err)
        if size == 1:
            def e2e_step(i=host_in, o=host_U):
                rc = L.pl_tsqr_svd_host_f64(o.data_ptr(), host_S.data_ptr(), host_V.data_ptr(), i.data_ptr(), m_e2e, n)
                _lib.check(rc, "pl_tsqr_svd_host_f64")
        else:
            # P ranks from host memory: ONE collective C call per rank (NCCL all-gather inside the library)
            comm = parall.c_comm()
            def e2e_step(i=host_in, o=host_U):
                rc = L.pl_tsqr_svd_host_dist_f64(comm, o.data_ptr(), host_S.data_ptr(), host_V.data_ptr(), i.data_ptr(), m_e2e, n)
                _lib.check(rc, "pl_tsqr_svd_host_dist_f64")
        t0 = time.perf_counter()
        e2e_step()
        barrier(cx)
        first_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier(cx)
        dt = max_over_ranks(cx, (time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": f_alg(m_e2e * size, n) / dt * 1e-9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": m_e2e * n * 8, "d2h_bytes_per_step": m_e2e * n * 8 + n * 8 + n * n * 8,
               "rows_per_gpu": m_e2e, "ms_per_step": dt * 1e3, "first_call_ms": first_ms,
               "host_buffers": "pinned (cudaHostAlloc via torch), first-touched on the GPU's NUMA node" if numa else "pinned",
               "numa_cpus": numa,
               "path": "pl_tsqr_svd_host_f64 (host pointers; row-chunk pipeline H2D || factor+Q, GEMM || D2H inside the timed region, "
                       "device buffers cached by the library after the first call, whose time is first_call_ms)" if size == 1
                       else "pl_tsqr_svd_host_dist_f64: ONE collective C call per rank (host pointers; chunked H2D || factor, ncclAllGather of the "
                            "n x n R inside the library, stack QR + Jacobi, chunked GEMM || D2H inside the timed region)"}
        if size == 1 and m_e2e == m:
            e2e["s_rel_diff_vs_device_path"] = float((host_S - S_dev).abs().max() / S_dev[0])
        if size == 1 and "pageable" not in skip:
            # the same call on plain (pageable) numpy arrays, as a numpy caller of the reference would pass them
            import numpy as np
            mp_rows = min(m_e2e, 2_000_000)
            a_np = host_in[:mp_rows].numpy().copy(); u_np = np.empty_like(a_np); s_np = np.empty(n); v_np = np.empty((n, n))
            def pg():
                _lib.check(L.pl_tsqr_svd_host_f64(u_np.ctypes.data, s_np.ctypes.data, v_np.ctypes.data, a_np.ctypes.data, mp_rows, n), "host")
            pg()
            t0 = time.perf_counter(); pg(); dtp = time.perf_counter() - t0
            e2e["pageable"] = {"value": f_alg(mp_rows, n) / dtp * 1e-9, "unit": "GFLOP/s", "rows": mp_rows, "ms_per_step": dtp * 1e3,
                               "note": "plain numpy (pageable) host arrays, same C call: the library stages them through its own pinned 64 MB ring with parallel host copies (PL_HOST_NO_STAGING=1 leaves it to cudaMemcpy's bounce buffer: 2.4 x slower)"}
            del a_np, u_np
    except Exception as ex:  # host memory too small etc.
        e2e = {"value": None, "unit": "GFLOP/s", "error": str(ex)[:300]}
    host_in = host_U = None
    L.pl_host_cache_free()
    release(cx)
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))
    except Exception:
        pass

    # ---- the other BASELINE shapes, per-GPU shards (driver-visible evidence for configs 1, 3, 4, 5 + strong scaling)
    other = {}
    if "other" not in skip:
        want = set(s for s in args.configs.split(",") if s)
        if "cfg5" in want:
            other["cfg5_shard_tsqr_svd_125000000x64"] = other_config(cx, "cfg5", 125_000_000, 64, "tsqr_svd", seed=2025)
            other["cfg5_shard_tsqr_svd_125000000x64"]["note"] = "1/8 of BASELINE configs[4] (1e9 x 64); with --gpus 8 the global matrix IS config 5"
        if "cfg3" in want:
            other["cfg3_shard_pod_24000000x256"] = other_config(cx, "cfg3", 24_000_000, 256, "POD.run(remove_mean=True)", seed=2023)
            other["cfg3_shard_pod_24000000x256"]["note"] = "1/8 of BASELINE configs[2] (64M points x 3 variables x 256 snapshots); with --gpus 8 the global matrix is config 3"
        if "cfg4" in want:
            other["cfg4_shard_dmd_2000000x1000"] = other_config(cx, "cfg4", 2_000_000, 1000, "DMD.run", steps=3, warmup=3, seed=2024)
            other["cfg4_shard_dmd_2000000x1000"]["note"] = "1/8 of BASELINE configs[3] (16M x 1000): tsqr_svd of the first 999 snapshots + U^T Y2 (all-reduce) + reduced eig + modes"
        if "cfg1" in want and size == 1:
            other["cfg1_pod_89351x151"] = other_config(cx, "cfg1", 89_351, 151, "POD.run(remove_mean=True)", steps=10, warmup=3, seed=2021)
        if "strong" in want and size > 1:
            sc = other_config(cx, "strong", ROWS_PER_GPU // size, n, "tsqr_svd", m_global=ROWS_PER_GPU, seed=SEED)
            sc["note"] = f"strong-scaling reading of configs[1]: the SAME global 8,000,000 x 512 matrix split over {size} GPUs"
            other["cfg2_strong_8000000x512"] = sc

    if rank == 0:
        cb = None
        if size == 1 and not args.no_cpu:
            try:
                cb = cpu_reference(1, args.cpu_rows if args.cpu_rows > 0 else 60_000, n, 2, 1)
            except Exception as ex:
                cb = None
                print(f"cpu baseline failed: {ex}", file=sys.stderr)
        value = f_alg(m_global, n) / (ms * 1e-3) * 1e-9
        line = {"metric": "tsqr_svd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": workload_config(args.rows, n, size, args.steps, args.warmup),
                "rows_snapshots_per_s": m_global * n / (ms * 1e-3),
                "frac_of_fp64_roofline": value * 1e-3 / (PEAK_FP64_TFLOPS * size),
                # transparency: the formulation executes 6 m n^2 (factor 2 + form Q 2 + GEMM 2) for n > 64
                "executed": {"flops": "6*m*n^2", "tflops": 1.5 * value * 1e-3, "frac_of_fp64_peak": 1.5 * value * 1e-3 / (PEAK_FP64_TFLOPS * size),
                             "ceiling_of_frac_of_fp64_roofline": 4.0 / 6.0},
                "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline, "phases_ms": phases,
                "checks": chk, "parity": parity, "other_configs": other}
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "ranks", "threads_per_rank")}
        print(json.dumps(line), flush=True)
    if size > 1:
        dist.barrier()
        parall.c_comm_destroy()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU")
    ap.add_argument("--cols", type=int, default=N_COLS)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0: chosen from steps + warmup)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per GPU of the e2e leg (0: the full shard when the host has room for the pinned buffers)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--skip", default="", help="comma list of legs to skip: parity,e2e,pageable,other,numa")
    ap.add_argument("--configs", default="cfg5,cfg3,cfg4,cfg1,strong", help="other_configs to time")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
