import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyloworder_b200 as pl
for n in (64, 256, 512, 999):
    A = torch.randn((4 * n, n), dtype=torch.float64, device="cuda")
    R = torch.linalg.qr(A, mode="r")[1]
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        U, S, V = pl.math.svd(R)
        torch.cuda.synchronize(); dt = time.time() - t0
    print("n", n, "svd ms", dt * 1e3, flush=True)
