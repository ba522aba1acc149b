"""CPU-only: pins the oracle against the reference's own known-answer vectors and checks
its two independent flavours against each other (see oracle/laser_oracle.h)."""
import numpy as np
import pytest

import oracle as O
from util import LAYOUTS, bf16_bits_to_f32, embed, extract, f32_to_bf16_bits, golden_cases

NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "i64"])
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_golden_numerics_faithful(case, dtype):
    a = np.array(case["a"], dtype=NP[dtype]); b = np.array(case["b"], dtype=NP[dtype])
    M, N, K = case["M"], case["N"], case["K"]
    c = np.full((M, N), 99, dtype=NP[dtype])
    O.gemm_strided(M, N, K, 1, a, K, 1, b, N, 1, 0, c, N, 1)
    assert np.array_equal(c, np.array(case["c"], dtype=NP[dtype]))


@pytest.mark.parametrize("isa", [1, 2, 3])
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_golden_structure_faithful(case, isa):
    if isa > O.detect_isa():
        pytest.skip("host lacks this ISA")
    a = np.array(case["a"], dtype=np.float32); b = np.array(case["b"], dtype=np.float32)
    M, N, K = case["M"], case["N"], case["K"]
    c = np.full((M, N), np.nan, dtype=np.float32)
    O.cpu_gemm_strided_f32(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1, isa)
    assert np.array_equal(c, np.array(case["c"], dtype=np.float32))


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_golden_bf16(case):
    a = f32_to_bf16_bits(np.array(case["a"], np.float32)); b = f32_to_bf16_bits(np.array(case["b"], np.float32))
    M, N, K = case["M"], case["N"], case["K"]
    c = np.zeros((M, N), np.uint16)
    O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1, bf16=True)
    assert np.array_equal(bf16_bits_to_f32(c), np.array(case["c"], np.float32))


SHAPES = [(1, 1, 1), (14, 32, 512), (15, 33, 513), (128, 128, 128), (200, 70, 1100), (193, 257, 31)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ab", [(1.0, 0.0), (0.5, -1.25), (1.0, 1.0), (2.0, 0.0)])
def test_two_flavours_bit_equal(shape, ab):
    """numerics-faithful chain == packed/micro-kernel/OpenMP restatement, bit for bit."""
    M, N, K = shape
    alpha, beta = ab
    a = O.fill_uniform_f32(M * K, 1, -0.1, 0.1); b = O.fill_uniform_f32(K * N, 2, -0.1, 0.1)
    c0 = O.fill_uniform_f32(M * N, 3, -1, 1)
    c_ref = c0.copy(); O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, c_ref, N, 1)
    for isa in range(1, O.detect_isa() + 1):
        if isa == 1 and M * N * K > 2e6:
            continue  # scalar kernel: keep the CPU suite fast
        if isa == 1:
            # the generic kernel is unfused (mul + add): only close, not bit-equal
            c = c0.copy(); O.cpu_gemm_strided_f32(M, N, K, alpha, a, K, 1, b, N, 1, beta, c, N, 1, isa)
            assert np.allclose(c, c_ref, rtol=1e-4, atol=1e-5)
            continue
        c = c0.copy(); O.cpu_gemm_strided_f32(M, N, K, alpha, a, K, 1, b, N, 1, beta, c, N, 1, isa)
        assert np.array_equal(c.view(np.uint32), c_ref.view(np.uint32)), (isa, shape, ab)


@pytest.mark.parametrize("la", LAYOUTS)
@pytest.mark.parametrize("lb", ["row", "col", "colslice"])
@pytest.mark.parametrize("lc", ["row", "col", "padded"])
def test_strided_views(la, lb, lc):
    M, N, K = 37, 29, 45
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32); B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    ba, oa, rsa, csa = embed(A, la); bb, ob, rsb, csb = embed(B, lb); bc, oc, rsc, csc = embed(C0, lc)
    want = C0.copy(); O.gemm_strided(M, N, K, 0.75, A, K, 1, B, N, 1, 0.5, want, N, 1)
    before = bc.copy()
    O.gemm_strided(M, N, K, 0.75, ba[oa:], rsa, csa, bb[ob:], rsb, csb, 0.5, bc[oc:], rsc, csc)
    assert np.array_equal(extract(bc, oc, rsc, csc, M, N), want)
    # nothing outside the C view was touched
    mask = np.ones(bc.size, bool); mask[(oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).ravel()] = False
    assert np.array_equal(bc[mask], before[mask])
    for isa in (2, 3):
        if isa > O.detect_isa():
            continue
        bc2, _, _, _ = embed(C0, lc)
        O.cpu_gemm_strided_f32(M, N, K, 0.75, ba[oa:], rsa, csa, bb[ob:], rsb, csb, 0.5, bc2[oc:], rsc, csc, isa)
        assert np.array_equal(bc2, bc)


def test_beta_zero_never_reads_c():
    M, N, K = 20, 40, 600
    a = O.fill_uniform_f32(M * K, 4, 0, 1); b = O.fill_uniform_f32(K * N, 5, 0, 1)
    clean = np.zeros(M * N, np.float32); O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, clean, N, 1)
    for fill in (np.nan, np.inf):
        c = np.full(M * N, fill, np.float32); O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1)
        assert np.array_equal(c, clean)
        c = np.full(M * N, fill, np.float32); O.cpu_gemm_strided_f32(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1)
        assert np.array_equal(c, clean)


def test_k_zero_leaves_c_untouched():
    c = np.arange(6, dtype=np.float32)
    a = np.zeros(1, np.float32)
    O.gemm_strided(2, 3, 0, 1.0, a, 0, 1, a, 3, 1, 0.5, c, 3, 1)  # gemm.nim:150: pc loop never runs
    assert np.array_equal(c, np.arange(6, dtype=np.float32))


def test_against_fp64_product():
    M, N, K = 96, 80, 4096
    a = O.fill_uniform_f32(M * K, 42, 0, 1); b = O.fill_uniform_f32(K * N, 43, 0, 1)
    c = np.zeros(M * N, np.float32); O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1)
    c64 = O.gemm_f32_in_f64(M, N, K, a, K, 1, b, N, 1).reshape(-1)
    assert O.max_relative_error(c, c64.astype(np.float32)) < 5e-6
    assert abs(c.astype(np.float64) - c64).max() / abs(c64).max() < 1e-6


def test_integer_wraparound():
    a = np.array([[2**31 - 1, 2]], dtype=np.int32); b = np.array([[2], [3]], dtype=np.int32)
    c = np.zeros((1, 1), np.int32)
    O.gemm_strided(1, 1, 2, 1, a, 2, 1, b, 1, 1, 0, c, 1, 1)
    assert c[0, 0] == np.int32((2 * (2**31 - 1) + 6) - 2**32)


def test_error_metrics():
    y = np.array([1.0, 2.0, 0.0, 4.0], np.float32); t = np.array([1.0, 2.2, 0.0, 2.0], np.float32)
    # error_functions.nim:6-26: |t-y| / max(|t|,|y|), 0 when both are 0
    want = (0 + 0.2 / 2.2 + 0 + 2.0 / 4.0) / 4
    assert abs(O.mean_relative_error(y, t) - want) < 1e-7
    assert abs(O.max_relative_error(y, t) - 1.0) < 1e-7
    assert O.normwise_relative_error(t, t) == 0.0


def test_fill_is_deterministic_and_in_range():
    x = O.fill_uniform_f32(10000, 42, -0.1, 0.1); y = O.fill_uniform_f32(10000, 42, -0.1, 0.1)
    assert np.array_equal(x, y) and x.min() >= -0.1 and x.max() < 0.1
    assert abs(float(x.mean())) < 5e-3 and not np.array_equal(x, O.fill_uniform_f32(10000, 43, -0.1, 0.1))
