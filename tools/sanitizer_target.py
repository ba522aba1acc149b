"""Small workload for compute-sanitizer: every kernel family once, modest sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laser_b200 as L, oracle as O
torch.cuda.set_device(0); L.init()
M, N, K = 300, 520, 400
A = O.fill_uniform_f32(M * K, 1, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 2, 0, 1).reshape(K, N)
want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
tA, tB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
tAt = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
for path in (L.PATH_SIMT, L.PATH_F16X3, L.PATH_TF32X3, L.PATH_TF32X1):
    tC = torch.zeros(M, N, device="cuda")
    L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, tC, N, 1, path=path)
    L.gemm_strided(M, N, K, 1.0, tAt, 1, M, tB, N, 1, 0.0, tC, N, 1, path=path)     # MN-major A, pair kernel
    torch.cuda.synchronize()
    print(L.PATH_NAMES[path], O.max_relative_error(tC.cpu().numpy(), want))
tC = torch.zeros(100, N, device="cuda")
L.gemm_strided(100, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, tC, N, 1, path=L.PATH_F16X3)   # single-CTA kernel
tC = torch.zeros(256, 256, device="cuda"); a2 = torch.rand(256, 4096, device="cuda"); b2 = torch.rand(4096, 256, device="cuda")
L.gemm_strided(256, 256, 4096, 1.0, a2, 4096, 1, b2, 256, 1, 0.0, tC, 256, 1)                # split-K + reduce
y = torch.zeros(2000, 3, device="cuda"); v = torch.rand(K, 3, device="cuda"); big = torch.rand(2000, K, device="cuda")
L.gemm_strided(2000, 3, K, 1.0, big, K, 1, v, 3, 1, 0.0, y, 3, 1)                            # GEMV
torch.cuda.synchronize(); print("done")
