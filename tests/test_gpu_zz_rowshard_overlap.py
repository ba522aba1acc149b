"""GPU (one device is enough): the overlapped variant of the row-shard driver (A prepared on a side stream while
the broadcast of B would be in flight; off by default, LASER_B200_ROWSHARD_OVERLAP=1).  Its first run on a B200 is
the round-end run of this file, which is why it sorts last among the GPU test files."""
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_rowsharded_overlap_prepack_single_gpu():
    """The overlapped variant (A prepared on a side stream while B would be in flight) against the
    plain call and the oracle; one GPU, no process group: only the stream choreography is new."""
    import numpy as np
    import oracle as O
    from laser_b200.rowshard import gemm_rowsharded

    M, N, K = 384, 520, 1000
    rng = np.random.default_rng(5)
    a = rng.random((M, K), dtype=np.float32)
    b = rng.random((K, N), dtype=np.float32)
    c0 = rng.random((M, N), dtype=np.float32)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    for alpha, beta in ((1.0, 0.0), (0.5, -1.25)):
        C1, C2 = torch.from_numpy(c0).cuda(), torch.from_numpy(c0).cuda()
        for _ in range(2):   # second round reuses the cached packed buffers
            C1.copy_(torch.from_numpy(c0))
            C2.copy_(torch.from_numpy(c0))
            gemm_rowsharded(M, N, K, alpha, A, B, beta, C1, broadcast=False, overlap_prepack=False)
            gemm_rowsharded(M, N, K, alpha, A, B, beta, C2, broadcast=False, overlap_prepack=True)
        torch.cuda.synchronize()
        ref = c0.copy()
        O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, ref, N, 1)
        got1, got2 = C1.cpu().numpy(), C2.cpu().numpy()
        scale = np.abs(ref).max()
        assert np.abs(got2 - ref).max() / scale < 1e-4
        assert np.abs(got2 - got1).max() / scale < 2e-6
