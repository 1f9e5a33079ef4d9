"""One POD.run + reconstruct at a mid size, for ncu captures of the streaming / panel / GEMM kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyloworder_b200 as pl
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 256
X = torch.rand((m, n), dtype=torch.float64, device="cuda")
mean = pl.math.temporal_mean(X)
Y = pl.math.subtract_mean(X, mean)
del Y
U, S, V = pl.POD.run(X, remove_mean=True)
Xr = pl.POD.reconstruct(*pl.POD.truncate(U, S, V, r=16))
torch.cuda.synchronize()
