"""Accuracy (vs the CPU oracle) and speed of the fp32-faithful path for the current LASER_B200_KC."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laser_b200 as L, oracle as O
torch.cuda.set_device(0); L.init()
kc = os.environ.get("LASER_B200_KC", "default")
for (M, N, K) in ((256, 512, 513), (256, 512, 4096), (129, 257, 8200), (512, 512, 16384)):
    for name, lo, hi in (("P", 0.0, 1.0), ("S", -0.1, 0.1)):
        A = O.fill_uniform_f32(M * K, 42, lo, hi).reshape(M, K); B = O.fill_uniform_f32(K * N, 43, lo, hi).reshape(K, N)
        want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
        exact = O.gemm_f32_in_f64(M, N, K, A, K, 1, B, N, 1)
        tA, tB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(); tC = torch.empty(M, N, device="cuda")
        res = {}
        for pname, path in (("x3", L.PATH_TF32X3), ("x1", L.PATH_TF32X1), ("f3", L.PATH_F16X3)):
            L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, tC, N, 1, path=path); torch.cuda.synchronize()
            got = tC.cpu().numpy()
            res[pname] = (O.max_relative_error(got, want), O.normwise_relative_error(got, want), O.mean_relative_error(got, want),
                          float(np.linalg.norm(got - exact) / np.linalg.norm(exact)), float(np.mean((got - exact) / np.abs(exact).mean())))
        ref_vs_exact = float(np.linalg.norm(want - exact) / np.linalg.norm(exact))
        print("kc=%s %dx%dx%d %s | x3: max %.2e norm %.2e mre %.2e vs_exact %.2e bias %+.2e | x1: max %.2e norm %.2e | f16x3: max %.2e norm %.2e mre %.2e vs_exact %.2e bias %+.2e | cpu_ref_vs_exact %.2e"
              % (kc, M, N, K, name, *res["x3"], res["x1"][0], res["x1"][1], *res["f3"], ref_vs_exact))
n = 8192
a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
for pname, path in (("x3", L.PATH_TF32X3), ("x1", L.PATH_TF32X1), ("f16x3", L.PATH_F16X3)):
    for _ in range(2): L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=path)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=path); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("kc=%s 8192^3 %s best %.3f ms %.1f TFLOP/s" % (kc, pname, min(ts), 2 * n**3 / min(ts) / 1e9))
