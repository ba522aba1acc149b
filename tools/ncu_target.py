"""Small ncu target: a few launches of each tensor-core kernel at the metric shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = int(os.environ.get("NCU_N", "8192"))
a = torch.empty(n * n, device="cuda"); b = torch.empty(n * n, device="cuda"); c = torch.empty(n * n, device="cuda")
L.fill_uniform_f32(a, n * n, 42, -0.1, 0.1); L.fill_uniform_f32(b, n * n, 43, -0.1, 0.1)
for _ in range(2):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1)                      # default: tf32 + bf16 cross terms
for _ in range(1):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=L.PATH_TF32X3)
for _ in range(1):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=L.PATH_TF32X1)
ab = a.view(n, n).to(torch.bfloat16); bb = b.view(n, n).to(torch.bfloat16); cb = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    L.gemm_strided(n, n, n, 1.0, ab, n, 1, bb, n, 1, 0.0, cb, n, 1)
torch.cuda.synchronize()
