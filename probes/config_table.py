"""Per-config timings on one B200 (BASELINE.json configs, per-GPU shards where the full config needs 8 GPUs)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyloworder_b200 as pl
from pyloworder_b200 import _lib, _dev
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
L = _lib.lib()
cases = [("cfg1 POD(remove_mean) 89,351 x 151 (full config)", 89351, 151, True, 2021),
         ("cfg3 shard POD(remove_mean) 24,000,000 x 256 (1/8 of 192M rows)", 24_000_000, 256, True, 2023),
         ("cfg4 shard tsqr_svd 2,000,000 x 999 (1/8 of 16M rows)", 2_000_000, 999, False, 2024),
         ("cfg5 shard tsqr_svd 125,000,000 x 64 (1/8 of 1e9 rows, in-place variant: A 64 GB + U 64 GB + 20 GB workspace)", 125_000_000, 64, False, 2025)]
out = []
for name, m, n, pod, seed in cases:
    try:
        X = bench.device_snapshots(torch, m, n, seed, 0, m, torch.device("cuda"))
        fn = (lambda: pl.POD.run(X, remove_mean=True)) if pod else (lambda: pl.math.tsqr_svd(X))
        for _ in range(2):
            U, S, V = fn(); del U
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            U, S, V = fn(); del U
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        del S, V
        U, S, V = fn()
        orth = float((U[:200000].T @ U[:200000]).diagonal().sub(0).abs().max())
        I = torch.eye(n, dtype=torch.float64, device="cuda")
        G = U.T @ U
        rec = {"config": name, "m": m, "n": n, "ms": round(ms, 2), "tflops_alg": round(4.0 * m * n * n / ms * 1e-9, 2),
               "frac_fp64_roofline": round(4.0 * m * n * n / ms * 1e-9 / 35.46, 3),
               "rows_snapshots_per_s": m * n / (ms * 1e-3), "UtU_minus_I_max": float((G - I).abs().max()),
               "VVt_minus_I_max": float((V @ V.T - I).abs().max())}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del X, U, S, V, G
    except Exception as e:
        print(json.dumps({"config": name, "error": str(e)[:300]}), flush=True)
    _dev.free_workspaces(); torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/r01_config_table.json", "w"), indent=1)
