"""Round-2 probe: f16x3 timings per shape / layout (device-resident, prep included) -> lines on stdout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, laser_b200 as L
torch.cuda.set_device(0); L.init()

def timeit(fn, iters=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

PATH = {"f16x3": L.PATH_F16X3, "tf32x3": L.PATH_TF32X3, "tf32x1": L.PATH_TF32X1}
for n in (4096, 8192):
    a = torch.rand(n, n, device="cuda") - 0.5; b = torch.rand(n, n, device="cuda") - 0.5; c = torch.empty(n, n, device="cuda")
    at = a.t().contiguous(); bt = b.t().contiguous()
    for name in ("f16x3", "tf32x3", "tf32x1"):
        p = PATH[name]
        for lay, fn in (("A row, B row", lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=p)),
                        ("A^T(col), B row", lambda: L.gemm_strided(n, n, n, 1.0, at, 1, n, b, n, 1, 0.0, c, n, 1, path=p)),
                        ("A row, B^T(col)", lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, bt, 1, n, 0.0, c, n, 1, path=p)),
                        ("A^T, B^T", lambda: L.gemm_strided(n, n, n, 1.0, at, 1, n, bt, 1, n, 0.0, c, n, 1, path=p))):
            ms = timeit(fn)
            L.profile_begin()
            for _ in range(3): fn()
            pr = L.profile_end()
            print("n=%d %-7s %-16s %.3f ms %.1f TFLOP/s | kernel %.3f ms prep %.3f ms" % (
                n, name, lay, ms, 2 * n**3 / ms / 1e9, pr["gemm_ms"] / max(1, pr["gemm_launches"]), pr["prep_ms"] / 3), flush=True)
# M=32768 single GPU (strong-scaling baseline)
n = 8192; m = 32768
a = torch.rand(m, n, device="cuda") - 0.5; b = torch.rand(n, n, device="cuda") - 0.5; c = torch.empty(m, n, device="cuda")
for name in ("f16x3",):
    ms = timeit(lambda: L.gemm_strided(m, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=PATH[name]), iters=4, warm=2)
    print("M=32768 N=K=8192 %-7s %.3f ms %.1f TFLOP/s" % (name, ms, 2.0 * m * n * n / ms / 1e9), flush=True)
