import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
n = 8192
a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.zeros(n, n, device="cuda")
ac = a[:, :4096].contiguous(); 
def t(fn, it=6):
    fn(); fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
for path, nm in ((L.PATH_F16X3, "f16x3"), (L.PATH_TF32X1, "x1"), (L.PATH_TF32X3, "x3")):
    print(nm, "full K=8192 beta=0        %.3f ms" % t(lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=path)))
    print(nm, "full K=8192 beta=1        %.3f ms" % t(lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 1.0, c, n, 1, path=path)))
    print(nm, "view K=4096 beta=0        %.3f ms" % t(lambda: L.gemm_strided(n, n, 4096, 1.0, a[:, :4096], n, 1, b[:4096], n, 1, 0.0, c, n, 1, path=path)))
    print(nm, "view K=4096 (2nd half) b=1 %.3f ms" % t(lambda: L.gemm_strided(n, n, 4096, 1.0, a[:, 4096:], n, 1, b[4096:], n, 1, 1.0, c, n, 1, path=path)))
    print(nm, "contig K=4096 beta=0      %.3f ms" % t(lambda: L.gemm_strided(n, n, 4096, 1.0, ac, 4096, 1, b[:4096], n, 1, 0.0, c, n, 1, path=path)))
    print(nm, "view K=1024 beta=1        %.3f ms" % t(lambda: L.gemm_strided(n, n, 1024, 1.0, a[:, 1024:2048], n, 1, b[1024:2048], n, 1, 1.0, c, n, 1, path=path)))
    L.profile_begin(); L.gemm_strided(n, n, 1024, 1.0, a[:, 1024:2048], n, 1, b[1024:2048], n, 1, 1.0, c, n, 1, path=path); print(nm, "K=1024 profile:", L.profile_end())
