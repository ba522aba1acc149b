"""ncu target: the default fp32 path at n^3 in the four operand layouts (A / A^T x B / B^T), one call each after a warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = int(os.environ.get("NCU_N", "8192"))
a = torch.rand(n, n, device="cuda") - 0.5; b = torch.rand(n, n, device="cuda") - 0.5; c = torch.empty(n, n, device="cuda")
at = a.t().contiguous(); bt = b.t().contiguous()
for rep in range(2):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1)      # A row, B row
    L.gemm_strided(n, n, n, 1.0, at, 1, n, b, n, 1, 0.0, c, n, 1)     # A^T (col), B row
    L.gemm_strided(n, n, n, 1.0, a, n, 1, bt, 1, n, 0.0, c, n, 1)     # A row, B^T (col)
    L.gemm_strided(n, n, n, 1.0, at, 1, n, bt, 1, n, 0.0, c, n, 1)    # A^T, B^T
torch.cuda.synchronize()
