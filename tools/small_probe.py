import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
for (M, N, K) in ((128, 128, 128), (256, 256, 256), (384, 384, 384), (512, 512, 512), (768, 768, 768), (1024, 1024, 1024), (2048, 64, 2048), (64, 2048, 2048), (4096, 256, 256)):
    a = torch.rand(M, K, device="cuda"); b = torch.rand(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
    ts = t(lambda: L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1, path=L.PATH_SIMT))
    tt = t(lambda: L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1, path=L.PATH_F16X3))
    print("%5dx%5dx%5d  simt %7.1f us   tc(f16x3) %7.1f us   %s" % (M, N, K, ts, tt, "SIMT" if ts < tt else "TC"))
