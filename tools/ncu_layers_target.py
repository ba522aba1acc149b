"""ncu target for the layer kernels: transposition (8192^2 f32), NCHW->NHWC, im2col at the reference conv bench
geometry, the strided forEach body of the reference's iteration benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = 8192
src = torch.rand(n * n, device="cuda"); dst = torch.empty_like(src)
for _ in range(2):
    L.transpose2D_copy(dst, src, n, n)
N, C, H, W = 64, 64, 112, 112
x = torch.rand(N * C * H * W, device="cuda"); y = torch.empty_like(x)
for _ in range(2):
    L.nchw2nhwc(y, x, N, C, H, W)
ish, ksh, pad, st = (16, 3, 224, 224), (20, 3, 3, 3), (0, 0), (1, 1)
inp = torch.rand(ish, device="cuda"); ws = torch.empty(16 * L.im2col_workspace_size(ish, ksh, pad, st), device="cuda")
for _ in range(2):
    L.im2col(ws, inp, ish, ksh, pad, st, images=16)
a = L.toTensor(np.random.rand(4096, 4096), "f64"); b = L.toTensor(np.random.rand(4096, 4096), "f64").transpose()
c = L.toTensor(np.random.rand(4096, 4096), "f64").transpose(); o = L.newTensor([4096, 4096], "f64")
for _ in range(2):
    L.forEach("bench", o, a, b, c)
torch.cuda.synchronize()
