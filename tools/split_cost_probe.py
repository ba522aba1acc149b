"""One GPU: what splitting one SGEMM 8192^3 into two partial products costs (no communication): K halves accumulating into C
(beta = 1 on the second), column halves of B / C (strided views), row halves of A / C."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = 8192; h = n // 2
A = torch.rand(n, n, device="cuda") - 0.5; B = torch.rand(n, n, device="cuda") - 0.5; C = torch.empty(n, n, device="cuda")
def P(t, off): return L.DevPtr(t.data_ptr() + 4 * off, "f32")
def timeit(fn, it=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
def whole(): L.gemm_strided(n, n, n, 1.0, A, n, 1, B, n, 1, 0.0, C, n, 1)
def k_halves():
    L.gemm_strided(n, n, h, 1.0, A, n, 1, B, n, 1, 0.0, C, n, 1)
    L.gemm_strided(n, n, h, 1.0, P(A, h), n, 1, P(B, h * n), n, 1, 1.0, C, n, 1)
def n_halves():
    L.gemm_strided(n, h, n, 1.0, A, n, 1, B, n, 1, 0.0, C, n, 1)
    L.gemm_strided(n, h, n, 1.0, A, n, 1, P(B, h), n, 1, 0.0, P(C, h), n, 1)
def m_halves():
    L.gemm_strided(h, n, n, 1.0, A, n, 1, B, n, 1, 0.0, C, n, 1)
    L.gemm_strided(h, n, n, 1.0, P(A, h * n), n, 1, B, n, 1, 0.0, P(C, h * n), n, 1)
for name, fn in (("whole", whole), ("two K halves (beta = 1 on the second)", k_halves), ("two column halves", n_halves), ("two row halves", m_halves), ("whole", whole)):
    print("%-40s %.3f ms" % (name, timeit(fn)), flush=True)
