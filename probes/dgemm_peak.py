"""Measure cuBLAS FP64 GEMM throughput (roofline denominator P_fp64). Library call, probe only."""
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
    # sustained: loop 3 s
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    cnt = 0; e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(4): c = a @ b
        cnt += 4; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_sustained_tflops"] = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
# tall-skinny shape like ours
m, n = 2_000_000, 512
a = torch.randn(m, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
c = a @ b; torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): c = a @ b
e1.record(); torch.cuda.synchronize()
res["dgemm_tall_2Mx512x512_tflops"] = 5 * 2 * m * n * n / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
open("gpurun_out/dgemm_peak.json", "w").write(json.dumps(res, indent=1))
