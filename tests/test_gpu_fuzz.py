"""GPU: seeded random sweep over shapes, strides, offsets, (alpha, beta) and kernel families
through the device C entry; every case is checked against the oracle (bit-exact for SIMT,
1e-4 max-relative on positive data for the fp32-faithful tensor-core modes)."""
import numpy as np
import pytest

import oracle as O
from backend import dev, sync
from util import LAYOUTS, embed, extract

pytestmark = pytest.mark.gpu
import laser_b200 as L  # noqa: E402

PATHS = [L.PATH_SIMT, L.PATH_F16X3, L.PATH_TF32X3, L.PATH_AUTO]


def one_case(rng):
    M = int(rng.choice([1, 2, 7, 64, 127, 128, 129, 255, 256, 257, 300, 513]))
    N = int(rng.choice([1, 3, 5, 31, 128, 129, 255, 256, 257, 384, 520]))
    K = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 300, 1000]))
    la, lb, lc = (str(rng.choice(LAYOUTS)) for _ in range(3))
    alpha, beta = float(rng.choice([1.0, 0.5, -2.0])), float(rng.choice([0.0, 0.0, 1.0, -1.25]))
    path = int(rng.choice(PATHS))
    return M, N, K, la, lb, lc, alpha, beta, path


@pytest.mark.parametrize("seed", range(12))
def test_random_cases(seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(16):
        M, N, K, la, lb, lc, alpha, beta, path = one_case(rng)
        A = O.fill_uniform_f32(M * K, seed * 100 + 1, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, seed * 100 + 2, 0, 1).reshape(K, N)
        C0 = O.fill_uniform_f32(M * N, seed * 100 + 3, 0, 1).reshape(M, N)
        if beta == 0.0:
            C0 = np.full((M, N), np.nan, np.float32)
        want = C0.copy(); O.gemm_strided(M, N, K, alpha, A, K, 1, B, N, 1, beta, want, N, 1)
        ba, oa, rsa, csa = embed(A, la); bb, ob, rsb, csb = embed(B, lb); bc, oc, rsc, csc = embed(C0, lc)
        ta, tb, tc = (dev(x) for x in (ba, bb, bc))
        L.gemm_strided(M, N, K, alpha, L.DevPtr(ta.data_ptr() + 4 * oa, "f32"), rsa, csa, L.DevPtr(tb.data_ptr() + 4 * ob, "f32"),
                       rsb, csb, beta, L.DevPtr(tc.data_ptr() + 4 * oc, "f32"), rsc, csc, path=path)
        sync()
        after = tc.cpu().numpy()
        got = extract(after, oc, rsc, csc, M, N)
        tag = (M, N, K, la, lb, lc, alpha, beta, path)
        if L.last_path() == L.PATH_SIMT and not (path == L.PATH_AUTO and N <= 4 and M >= 1024):
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), tag
        else:
            # alpha*AB + beta*C can cancel (negative alpha or beta): gate on the magnitude of the terms
            scale = np.abs(alpha) * (A @ B) + np.abs(beta) * np.abs(np.nan_to_num(C0))
            assert np.max(np.abs(got - want) / scale) < 1e-4, tag
        mask = np.ones(after.size, bool); mask[(oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).ravel()] = False
        assert np.array_equal(after[mask], bc[mask], equal_nan=True), tag
