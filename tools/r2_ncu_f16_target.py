"""ncu target: the f16x3 path at the metric shape (abs-max + split + GEMM), row-major operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = int(os.environ.get("NCU_N", "8192"))
a = torch.empty(n * n, device="cuda"); b = torch.empty(n * n, device="cuda"); c = torch.empty(n * n, device="cuda")
L.fill_uniform_f32(a, n * n, 42, -0.1, 0.1); L.fill_uniform_f32(b, n * n, 43, -0.1, 0.1)
for _ in range(int(os.environ.get("NCU_REPS", "2"))):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=L.PATH_F16X3)
torch.cuda.synchronize()
