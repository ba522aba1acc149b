"""ncu target: the other kernels of round 2 once each -- fp64 DMMA, few-rows exact kernel (the im2col convolution), split-K
tail reduce (4096^3), im2col, transposes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = 4096
a = torch.rand(n, n, dtype=torch.float64, device="cuda"); b = torch.rand(n, n, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1)                      # gemm_dmma_kernel
af = torch.rand(n, n, device="cuda"); bf = torch.rand(n, n, device="cuda"); cf = torch.empty(n, n, device="cuda")
L.gemm_strided(n, n, n, 1.0, af, n, 1, bf, n, 1, 0.0, cf, n, 1)                   # gemm_tc_kernel + splitk_tail_reduce_kernel
ish, ksh, pad, st = (16, 3, 224, 224), (20, 3, 3, 3), (0, 0), (1, 1)
osh = L.conv2d_out_shape(ish, ksh, pad, st)
per = L.im2col_workspace_size(ish, ksh, pad, st)
inp = torch.rand(ish, device="cuda"); ker = torch.rand(ksh, device="cuda")
ws = torch.empty(ish[0] * per, device="cuda"); out = torch.empty(osh, device="cuda")
L.conv2d_im2col(out, inp, ish, ker, ksh, pad, st, workspace=ws, workspace_images=16)   # im2col_kernel + gemm_skinny_m_kernel
src = torch.rand(8192 * 8192, device="cuda"); dst = torch.empty_like(src)
L.transpose2D_copy(dst, src, 8192, 8192)                                           # transpose_batched_kernel
torch.cuda.synchronize()
