"""ncu target for the non-tensor-core kernels: exact SIMT GEMM, warp-shuffle GEMV, gather/pack."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = 4096
a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
for _ in range(2):
    L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=L.PATH_SIMT)
v = torch.rand(n, 3, device="cuda"); y = torch.empty(16384, 3, device="cuda"); big = torch.rand(16384, n, device="cuda")
for _ in range(2):
    L.gemm_strided(16384, 3, n, 1.0, big, n, 1, v, 3, 1, 0.0, y, 3, 1)           # skinny: gemv_warp_kernel
pb = L.alloc_packed(L.gemm_prepackB_mem_required(n, n, n))
for _ in range(2):
    L.gemm_prepackB(pb, n, n, n, b, n, 1)                                        # transposing gather + split
torch.cuda.synchronize()
