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
# ---- kernels of round 2
a3 = torch.rand(600, 2048, device="cuda"); b3 = torch.rand(2048, 512, device="cuda"); c3 = torch.zeros(600, 512, device="cuda")
L.gemm_strided(600, 512, 2048, 1.0, a3, 2048, 1, b3, 512, 1, 0.0, c3, 512, 1)                # ring-buffered preparation (rows of 2048 floats)
err = ((c3 - a3 @ b3).abs().max() / (a3 @ b3).abs().max()).item(); print("ring prep", err); assert err < 1e-4
a4 = torch.rand(2560, 2048, device="cuda"); b4 = torch.rand(2048, 2304, device="cuda"); c4 = torch.zeros(2560, 2304, device="cuda")
n0 = L.launch_count()
L.gemm_strided(2560, 2304, 2048, 1.0, a4, 2048, 1, b4, 2304, 1, 0.0, c4, 2304, 1)            # 90 pair-tiles on 74 pairs: split-K of the last wave
print("tail split launches", L.launch_count() - n0)
err = ((c4 - a4 @ b4).abs().max() / (a4 @ b4).abs().max()).item(); print("tail split", err); assert err < 1e-4
a5 = torch.rand(1152, 96, device="cuda", dtype=torch.float64); b5 = torch.rand(96, 1280, device="cuda", dtype=torch.float64)
c5 = torch.zeros(1152, 1280, device="cuda", dtype=torch.float64)
d0 = L.lib().laser_b200_debug_f64_dmma_launches()
L.gemm_strided(1152, 1280, 96, 1.0, a5, 96, 1, b5, 1280, 1, 0.0, c5, 1280, 1)                 # fp64 tensor cores (cp.async staging)
assert L.lib().laser_b200_debug_f64_dmma_launches() == d0 + 1
err = ((c5 - a5 @ b5).abs().max() / (a5 @ b5).abs().max()).item(); print("dmma", err); assert err < 1e-14
a6 = torch.rand(20, 27, device="cuda"); b6 = torch.rand(27, 5001, device="cuda"); c6 = torch.zeros(20, 5001, device="cuda")
L.gemm_strided(20, 5001, 27, 1.0, a6, 27, 1, b6, 5001, 1, 0.0, c6, 5001, 1, path=L.PATH_SIMT)  # few-rows kernel, ragged N
err = ((c6 - a6 @ b6).abs().max() / (a6 @ b6).abs().max()).item(); print("few rows", err); assert err < 1e-5
torch.cuda.synchronize(); print("done")
