"""GPU: fused epilogue C <- act(alpha*A*B + beta*C + bias) (the reference's stated next step,
gemm.nim:196), for the exact and the tensor-core kernel families, vectorised and strided C."""
import numpy as np
import pytest

import oracle as O
from backend import dev, sync

pytestmark = pytest.mark.gpu
import laser_b200 as L  # noqa: E402

ACTS = {"none": lambda x: x, "relu": lambda x: np.maximum(x, 0), "tanh": np.tanh,
        "sigmoid": lambda x: 1.0 / (1.0 + np.exp(-x))}


@pytest.mark.parametrize("act", list(ACTS))
@pytest.mark.parametrize("per_row", [False, True])
@pytest.mark.parametrize("path", [L.PATH_SIMT, L.PATH_F16X3, L.PATH_TF32X3])
@pytest.mark.parametrize("ldc", [520, 523])
def test_fused_bias_activation(act, per_row, path, ldc):
    M, N, K = 300, 520, 700
    A = O.fill_uniform_f32(M * K, 1, -0.1, 0.1).reshape(M, K); B = O.fill_uniform_f32(K * N, 2, -0.1, 0.1).reshape(K, N)
    C0 = O.fill_uniform_f32(M * N, 3, -1, 1).reshape(M, N)
    bias = O.fill_uniform_f32(M if per_row else N, 4, -1, 1)
    base = C0.copy(); O.gemm_strided(M, N, K, 0.5, A, K, 1, B, N, 1, -1.25, base, N, 1)
    want = ACTS[act]((base.astype(np.float64) + (bias[:, None] if per_row else bias[None, :]))).astype(np.float32)
    tA, tB = dev(A), dev(B)
    hbuf = np.full((M, ldc), -7.0, np.float32); hbuf[:, :N] = C0
    buf = dev(hbuf)
    L.gemm_strided_fused(M, N, K, 0.5, tA, K, 1, tB, N, 1, -1.25, buf, ldc, 1, bias=dev(bias),
                         bias_per_row=per_row, activation=act, path=path)
    sync()
    after = buf.cpu().numpy()
    got = after[:, :N]
    assert np.allclose(got, want, rtol=2e-5, atol=2e-6), np.abs(got - want).max()
    assert np.all(after[:, N:] == -7.0)
    # the thread-local epilogue must not leak into the next plain call
    tC = dev(C0); L.gemm_strided(M, N, K, 0.5, tA, K, 1, tB, N, 1, -1.25, tC, N, 1, path=path)
    sync()
    assert np.allclose(tC.cpu().numpy(), base, rtol=2e-5, atol=2e-6)
