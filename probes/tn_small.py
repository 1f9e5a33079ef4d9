"""Short driver for ncu: matmul_tn (X^T Y) at rows x a x b (argv)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloworder_b200.vmmath.maths import matmul_tn
m, a, b = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X = torch.randn((m, a), dtype=torch.float64, device="cuda")
Y = torch.randn((m, b), dtype=torch.float64, device="cuda")
for _ in range(2):
    C = matmul_tn(X, Y)
torch.cuda.synchronize()
print("done", float(C[0, 0]))
