"""Small ncu target: the two opt-in two-piece fp32 modes at the metric shape (preparation kernels + GEMM), after one
call of the default mode for reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = int(os.environ.get("NCU_N", "8192"))
a = torch.empty(n * n, device="cuda"); b = torch.empty(n * n, device="cuda"); c = torch.empty(n * n, device="cuda")
L.fill_uniform_f32(a, n * n, 42, -0.1, 0.1); L.fill_uniform_f32(b, n * n, 43, -0.1, 0.1)
L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1)                             # default: tf32 + bf16 cross terms
for path in (L.PATH_BF16X3, L.PATH_F16X3):
    for _ in range(2):
        L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=path)          # A K-major, B MN-major (per-column scales)
    L.gemm_strided(n, n, n, 1.0, a, 1, n, b, n, 1, 0.0, c, n, 1, path=path)              # A^T: MN-major A as well
torch.cuda.synchronize()
