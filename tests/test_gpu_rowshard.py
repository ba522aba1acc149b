"""GPU, >= 2 devices: the NCCL row-sharded path end to end (torchrun, one rank per GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("env", [{}, {"RS_COLMAJOR": "1"}, {"LASER_B200_ROWSHARD_PANELS": "0"}, {"LASER_B200_ROWSHARD_PANELS": "4", "RS_M": "5000"}],
                         ids=["prepared-panels", "column-major-B", "raw-broadcast", "four-panels-uneven-rows"])
def test_rowsharded_nccl(env):
    """every rank checks sampled rows of its C panel against the oracle: B prepared on the root and sent in column panels
    (row- and column-major B), and the raw broadcast of B"""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "rowshard_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=550, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("max_rel_err") == n


def test_host_entry_row_shards_over_the_gpus_of_this_process():
    """laser_b200_gemm_rowsharded_f32: host buffers, one process driving 2 (4) devices, NCCL group broadcast of B"""
    import numpy as np
    import oracle as O
    from laser_b200 import rowshard as RS

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for g in ([2, 4] if n >= 4 else [2]):
        for (M, N, K, alpha, beta) in ((4096, 1024, 1536, 1.0, 0.0), (3000, 520, 900, 0.5, -1.25)):
            a = O.fill_uniform_f32(M * K, 3, -1, 1).reshape(M, K); b = O.fill_uniform_f32(K * N, 4, -1, 1).reshape(K, N)
            c0 = O.fill_uniform_f32(M * N, 5, -1, 1).reshape(M, N)
            c = c0.copy() if beta else np.full((M, N), np.nan, np.float32)
            RS.gemm_rowsharded_host(g, M, N, K, alpha, a, K, 1, b, N, 1, beta, c, N, 1)
            want = c0.copy()
            O.cpu_gemm_strided_f32(M, N, K, alpha, a.reshape(-1), K, 1, b.reshape(-1), N, 1, beta, want.reshape(-1), N, 1)
            assert O.normwise_relative_error(c, want) < 2e-6, (g, M, N, K)


def test_one_process_two_devices():
    """A single process driving two GPUs through the C ABI (per-device context, tensor-map cache,
    function attributes): the second device must behave like the first."""
    import numpy as np
    import oracle as O
    import laser_b200 as L

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    M, N, K = 300, 260, 520
    a = O.fill_uniform_f32(M * K, 1, 0, 1).reshape(M, K); b = O.fill_uniform_f32(K * N, 2, 0, 1).reshape(K, N)
    ref = np.zeros((M, N), np.float32)
    O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, ref, N, 1)
    try:
        for d in (0, 1, 0):
            torch.cuda.set_device(d)
            A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
            C = torch.full((M, N), float("nan"), device="cuda")
            L.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, C, N, 1)                       # tensor cores
            E = torch.empty((M, N), device="cuda")
            L.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, E, N, 1, path=L.PATH_SIMT)      # exact kernel
            v = torch.empty((M, 1), device="cuda")
            L.gemm_strided(M, 1, K, 1.0, A, K, 1, B, N, 1, 0.0, v, 1, 1)                       # column of B (small: exact)
            torch.cuda.synchronize()
            assert C.device.index == d
            assert np.abs(C.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4
            assert np.array_equal(E.cpu().numpy(), ref)
    finally:
        torch.cuda.set_device(0)
