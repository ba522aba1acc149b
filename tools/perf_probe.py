"""Quick GPU probe (not the bench): TF32 operand rounding behaviour + raw kernel timings."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import laser_b200 as L

torch.cuda.set_device(0)
L.init()

def probe_rounding():
    # value v = 1 + 2^-11 + 2^-12: tf32 truncation -> 1.0, round-to-nearest -> 1 + 2^-10
    # both operands K-major (A row-major, B given as B^T storage) so that TMA feeds the raw bits
    M, N, K = 128, 256, 32
    for name, v in (("1+2^-11+2^-12", 1 + 2**-11 + 2**-12), ("1+2^-11 (tie)", 1 + 2**-11), ("1+2^-10", 1 + 2**-10)):
        x = torch.full((M, K), v, dtype=torch.float32, device="cuda"); onesT = torch.ones((N, K), dtype=torch.float32, device="cuda")
        c = torch.zeros((M, N), dtype=torch.float32, device="cuda")
        L.gemm_strided(M, N, K, 1.0, x, K, 1, onesT, 1, K, 0.0, c, N, 1, path=L.PATH_TF32X1)
        torch.cuda.synchronize()
        print("rounding probe A=%s: C/K = %.10f  (trunc -> 1.0, rn -> %.10f)" % (name, c[0, 0].item() / K, 1 + 2**-10))
        xT = torch.full((N, K), v, dtype=torch.float32, device="cuda"); ones = torch.ones((M, K), dtype=torch.float32, device="cuda")
        c.zero_()
        L.gemm_strided(M, N, K, 1.0, ones, K, 1, xT, 1, K, 0.0, c, N, 1, path=L.PATH_TF32X1)
        torch.cuda.synchronize()
        print("               B=%s: C/K = %.10f" % (name, c[0, 0].item() / K))

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

def perf():
    out = {}
    for n in (4096, 8192):
        a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
        at = a.t().contiguous()
        for pname, path in (("tf32x1", L.PATH_TF32X1), ("tf32x3", L.PATH_TF32X3)) + ((("simt", L.PATH_SIMT),) if n == 4096 else ()):
            mn, av = timeit(lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=path), iters=3 if pname == "simt" else 5)
            print("n=%d %-7s row,row     best %.3f ms  %.1f TFLOP/s (avg %.3f ms)" % (n, pname, mn, 2 * n**3 / mn / 1e9, av)); out["%d_%s" % (n, pname)] = mn
        for pname, path in (("tf32x1", L.PATH_TF32X1), ("tf32x3", L.PATH_TF32X3)):
            mn, av = timeit(lambda: L.gemm_strided(n, n, n, 1.0, at, 1, n, b, n, 1, 0.0, c, n, 1, path=path))
            print("n=%d %-7s A^T(col),row best %.3f ms  %.1f TFLOP/s" % (n, pname, mn, 2 * n**3 / mn / 1e9))
            bt = b.t().contiguous()
            mn, av = timeit(lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, bt, 1, n, 0.0, c, n, 1, path=path))
            print("n=%d %-7s row,B^T(col) best %.3f ms  %.1f TFLOP/s  (both K-major)" % (n, pname, mn, 2 * n**3 / mn / 1e9))
        ab = a.to(torch.bfloat16); bb = b.to(torch.bfloat16); cb = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
        mn, av = timeit(lambda: L.gemm_strided(n, n, n, 1.0, ab, n, 1, bb, n, 1, 0.0, cb, n, 1))
        print("n=%d bf16   row,row      best %.3f ms  %.1f TFLOP/s" % (n, mn, 2 * n**3 / mn / 1e9))
        torch.backends.cuda.matmul.allow_tf32 = True
        mn, av = timeit(lambda: torch.matmul(a, b, out=c))
        print("n=%d cuBLAS tf32 (informal ceiling) best %.3f ms %.1f TFLOP/s" % (n, mn, 2 * n**3 / mn / 1e9))
        mn, av = timeit(lambda: torch.matmul(ab, bb, out=cb))
        print("n=%d cuBLAS bf16 (informal ceiling) best %.3f ms %.1f TFLOP/s" % (n, mn, 2 * n**3 / mn / 1e9))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/perf_probe.json", "w"))

def perf_packed():
    n = 8192
    a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
    pa = L.alloc_packed(L.gemm_prepackA_mem_required(n, n, n)); pb = L.alloc_packed(L.gemm_prepackB_mem_required(n, n, n))
    mn, _ = timeit(lambda: L.gemm_prepackB(pb, n, n, n, b, n, 1)); print("prepackB 8192^2 (transposing gather) best %.3f ms" % mn)
    mn, _ = timeit(lambda: L.gemm_prepackA(pa, n, n, n, a, n, 1)); print("prepackA 8192^2 best %.3f ms" % mn)
    for name, fn in (("default (split every call)", lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1)),
                     ("packedB (A split per call)", lambda: L.gemm_packedB(n, n, n, 1.0, a, n, 1, pb, 0.0, c, n, 1)),
                     ("packed A and B", lambda: L.gemm_packed(n, n, n, 1.0, pa, pb, 0.0, c, n, 1))):
        mn, av = timeit(fn, iters=8)
        print("n=8192 %-28s best %.3f ms  %.1f TFLOP/s (avg %.3f)" % (name, mn, 2 * n**3 / mn / 1e9, av))

def perf_midsize():
    for (M, N, K) in ((512, 512, 512), (1024, 1024, 1024), (1920, 1920, 1920), (1024, 1024, 8192), (256, 256, 65536), (2048, 2048, 2048), (4096, 4096, 4096)):
        a = torch.rand(M, K, device="cuda"); b = torch.rand(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
        mn, av = timeit(lambda: L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1), iters=10)
        print("midsize %dx%dx%d default best %.1f us  %.1f TFLOP/s" % (M, N, K, mn * 1e3, 2.0 * M * N * K / mn / 1e9))

def perf_pcie():
    n = 8192
    h = torch.empty(n * n, dtype=torch.float32).pin_memory(); d = torch.empty(n * n, device="cuda")
    mn, _ = timeit(lambda: d.copy_(h, non_blocking=True)); print("pinned H2D 268 MB: %.2f ms  %.1f GB/s" % (mn, 0.268435 / mn * 1e3))
    mn, _ = timeit(lambda: h.copy_(d, non_blocking=True)); print("pinned D2H 268 MB: %.2f ms  %.1f GB/s" % (mn, 0.268435 / mn * 1e3))

if __name__ == "__main__":
    probe_rounding()
    perf_pcie()
    perf_midsize()
    perf_packed()
    perf()
