"""CPU-only: the exact CUDA-core GEMM kernel (laser_b200/csrc/gemm_simt.cuh: gemm_simt_kernel, the
path that has to be bit-identical to the reference's CPU order of operations) executed on host
threads (tests/emu/) with the library's own launch planning, against the oracle: the reference's
known-answer vectors for every dtype, strided layouts, alpha/beta, kc-block boundaries, wrapping
integer arithmetic, the fused epilogue."""
import ctypes

import numpy as np
import pytest

import oracle as O
from emu_build import build_emu
from util import LAYOUTS, embed, extract, golden_cases

i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
SCALAR = {"f32": ctypes.c_float, "f64": ctypes.c_double, "i32": ctypes.c_int32, "i64": ctypes.c_int64}
NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}


@pytest.fixture(scope="module")
def emu():
    L = ctypes.CDLL(build_emu("simt_emu", ["gemm_simt.cuh", "gemm_simt_kernel.inc", "gemm_dmma.cuh", "ptx.cuh"]))
    f64 = ctypes.c_double
    L.emu_gemm_skinny_m_f32.restype = ci
    L.emu_gemm_skinny_m_f32.argtypes = [i64, i64, i64, i64, ctypes.c_float, vp, i64, i64, i64, vp, i64, i64, i64, ctypes.c_float, vp,
                                        i64, i64, i64, ci, vp, ci, ci]
    L.emu_gemm_dmma_f64.restype = ci
    L.emu_gemm_dmma_f64.argtypes = [i64, i64, i64, i64, f64, vp, i64, i64, i64, vp, i64, i64, i64, f64, vp, i64, i64, i64, ci]
    for name, sc in SCALAR.items():
        fn = getattr(L, "emu_gemm_simt_" + name)
        fn.restype = ci
        fn.argtypes = [i64, i64, i64, sc, vp, i64, i64, vp, i64, i64, sc, vp, i64, i64, ci] + \
                      ([vp, ci, ci] if name == "f32" else [])
    f32 = ctypes.c_float
    L.emu_gemm_simt_batched_f32.restype = ci
    L.emu_gemm_simt_batched_f32.argtypes = [i64, i64, i64, i64, f32, vp, i64, i64, i64, vp, i64, i64, i64, f32, vp, i64, i64,
                                            i64, ci]
    for name, sc in (("f64", ctypes.c_double), ("i64", i64)):
        fn = getattr(L, "emu_gemm_simt_batched_" + name)
        fn.restype = ci
        fn.argtypes = [i64, i64, i64, i64, sc, vp, i64, i64, i64, vp, i64, i64, i64, sc, vp, i64, i64, i64, ci]
    L.emu_gemv.restype = ci
    L.emu_gemv.argtypes = [ci, ci, i64, i64, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, i64, ci]
    return L


def at(buf, off):
    return ctypes.c_void_p(buf.ctypes.data + off * buf.itemsize)


def run(emu, name, M, N, K, alpha, A, oa, rsa, csa, B, ob, rsb, csb, beta, C, oc, rsc, csc, grid=0, epi=None):
    extra = []
    if name == "f32":
        bias, per_row, act = epi if epi else (None, 0, 0)
        extra = [ctypes.c_void_p(bias.ctypes.data) if bias is not None else None, per_row, act]
    return getattr(emu, "emu_gemm_simt_" + name)(M, N, K, alpha, at(A, oa), rsa, csa, at(B, ob), rsb, csb, beta,
                                                 at(C, oc), rsc, csc, grid, *extra)


@pytest.mark.parametrize("name", ["f32", "f64", "i32", "i64"])
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_known_answer_vectors(emu, name, case):
    dt = NP[name]
    M, N, K = case["M"], case["N"], case["K"]
    a = np.array(case["a"], dt); b = np.array(case["b"], dt)
    c = np.full((M, N), 99, dt)
    run(emu, name, M, N, K, 1, a, 0, K, 1, b, 0, N, 1, 0, c, 0, N, 1)
    assert np.array_equal(c, np.array(case["c"], dt))


@pytest.mark.parametrize("la,lb,lc", [(a, b, c) for a in LAYOUTS for b, c in (("row", "row"), ("col", "both2"))] +
                         [("row", b, "negrow") for b in LAYOUTS] + [("colslice", "padded", c) for c in LAYOUTS])
def test_strided_layouts_bit_exact(emu, la, lb, lc):
    M, N, K = 70, 45, 90
    rng = np.random.default_rng(1)
    a = rng.random((M, K), dtype=np.float32); b = rng.random((K, N), dtype=np.float32)
    c0 = rng.random((M, N), dtype=np.float32)
    A, oa, rsa, csa = embed(a, la); B, ob, rsb, csb = embed(b, lb); C, oc, rsc, csc = embed(c0, lc)
    Cref = C.copy()
    O.gemm_strided(M, N, K, 1.0, A[oa:] if oa >= 0 else A, rsa, csa, B[ob:], rsb, csb, 2.0, Cref[oc:], rsc, csc)
    run(emu, "f32", M, N, K, 1.0, A, oa, rsa, csa, B, ob, rsb, csb, 2.0, C, oc, rsc, csc, grid=2)
    assert np.array_equal(C, Cref)          # the whole buffer: nothing outside the view is touched


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (128, 128, 128), (129, 127, 513), (5, 300, 1030), (257, 3, 17)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (1.0, 1.0), (0.5, -1.25)])
def test_shapes_and_scalars(emu, M, N, K, alpha, beta):
    a = O.fill_uniform_f32(M * K, 5, -0.1, 0.1).reshape(M, K); b = O.fill_uniform_f32(K * N, 6, -0.1, 0.1).reshape(K, N)
    c = O.fill_uniform_f32(M * N, 7, -1, 1).reshape(M, N)
    if beta == 0.0:
        c[:] = np.nan                        # beta == 0: C is not read (gemm_ukernel_generic.nim:60-62)
    ref = c.copy()
    O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, ref, N, 1)
    run(emu, "f32", M, N, K, alpha, a, 0, K, 1, b, 0, N, 1, beta, c, 0, N, 1, grid=3)
    if alpha == 1.0:
        assert np.array_equal(c, ref)        # bit-identical, across kc = 512 block boundaries too
    else:                                    # alpha != 1: the oracle's compiler may contract C += alpha*AB (1 ulp)
        assert np.abs(c - ref).max() <= 2e-7 * np.abs(ref).max()


def test_f64_and_integers(emu):
    M, N, K = 40, 50, 300                    # kc = 256 for 8-byte types
    rng = np.random.default_rng(2)
    a = rng.random((M, K)); b = rng.random((K, N)); c = rng.random((M, N)); ref = c.copy()
    O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 1.0, ref, N, 1)
    run(emu, "f64", M, N, K, 1.0, a, 0, K, 1, b, 0, N, 1, 1.0, c, 0, N, 1)
    assert np.array_equal(c, ref)
    for name, big in (("i32", 2**31 - 5), ("i64", 2**63 - 5)):
        dt = NP[name]
        ai = rng.integers(-big, big, size=(M, K), dtype=dt); bi = rng.integers(-big, big, size=(K, N), dtype=dt)
        ci_ = rng.integers(-9, 9, size=(M, N), dtype=dt); refi = ci_.copy()
        O.gemm_strided(M, N, K, 3, ai, K, 1, bi, N, 1, -2, refi, N, 1)
        run(emu, name, M, N, K, 3, ai, 0, K, 1, bi, 0, N, 1, -2, ci_, 0, N, 1, grid=1)
        assert np.array_equal(ci_, refi)     # wrapping arithmetic, as the reference's mullo + add


@pytest.mark.parametrize("M,N,K,alpha,beta,layout", [(130, 140, 300, 1.0, 0.0, "nn"), (200, 129, 515, 1.0, 1.0, "tn"),
                                                     (64, 260, 257, 1.0, -0.5, "nt"), (129, 130, 40, 2.0, 0.0, "tt"),
                                                     (5, 7, 3, 1.0, 1.0, "nn")])
def test_f64_tensor_core_kernel_is_the_same_fma_chain(emu, M, N, K, alpha, beta, layout):
    """gemm_dmma.cuh on the emulator's model of mma.sync.m8n8k4.f64 (per element: four fma steps in k order): bit-identical
    to the oracle across kc = 256 block boundaries, ragged tiles and all operand layouts; C beyond the view untouched"""
    rng = np.random.default_rng(5)
    a = rng.standard_normal((M, K)); b = rng.standard_normal((K, N)); c0 = rng.standard_normal((M, N + 3))
    A, rsa, csa = (a, K, 1) if layout[0] == "n" else (np.ascontiguousarray(a.T), 1, M)
    B, rsb, csb = (b, N, 1) if layout[1] == "n" else (np.ascontiguousarray(b.T), 1, K)
    ref = c0.copy(); O.gemm_strided(M, N, K, alpha, A, rsa, csa, B, rsb, csb, beta, ref, N + 3, 1)
    got = c0.copy()
    if beta == 0.0:
        got[:, :N] = np.nan
    tiles = emu.emu_gemm_dmma_f64(1, M, N, K, alpha, at(A, 0), rsa, csa, 0, at(B, 0), rsb, csb, 0, beta, at(got, 0), N + 3, 1, 0, 3)
    assert tiles == -(-M // 128) * -(-N // 128)
    if alpha == 1.0:
        assert np.array_equal(got, ref)
    else:                                    # alpha != 1: the oracle's compiler may contract C += alpha*AB (1 ulp)
        assert np.abs(got - ref).max() <= 4e-16 * np.abs(ref).max() and np.array_equal(got[:, N:], ref[:, N:])


@pytest.mark.parametrize("M,N,K,alpha,beta,b_col,batch", [(20, 2500, 27, 1.0, 0.0, False, 3), (7, 1030, 600, 1.0, 1.0, False, 1),
                                                          (16, 1100, 513, 0.5, -1.25, True, 2), (32, 1029, 70, 1.0, 0.0, False, 1),
                                                          (1, 2049, 5, 1.0, 1.0, True, 1), (24, 2052, 530, 0.5, 2.0, False, 2),
                                                          (9, 1500, 65, 1.0, 0.0, False, 1)])
def test_skinny_m_kernel_bit_exact(emu, M, N, K, alpha, beta, b_col, batch):
    """few output rows x wide N (the im2col convolution's GEMM): bit-identical to the oracle across kc = 512 blocks, shared A
    across the batch, row- and column-major B, ragged N (scalar tail next to the vector path), bias + relu after the last block"""
    rng = np.random.default_rng(8)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bl = rng.standard_normal((batch, K, N)).astype(np.float32)
    C0 = rng.standard_normal((batch, M, N)).astype(np.float32)
    bias = rng.standard_normal(M).astype(np.float32)
    ldb = -(-N // 4) * 4                     # row-major B is stored with a 16-byte multiple pitch (padding columns hold junk)
    if b_col:
        B = np.ascontiguousarray(Bl.transpose(0, 2, 1)); rsb, csb, bsb = 1, K, K * N
    else:
        B = np.full((batch, K, ldb), 7e30, np.float32); B[:, :, :N] = Bl; rsb, csb, bsb = ldb, 1, K * ldb
    ref = C0.copy()
    for i in range(batch):
        O.gemm_strided(M, N, K, alpha, A, K, 1, B[i], rsb, csb, beta, ref[i], N, 1)
    ref_epi = np.maximum(ref + bias[None, :, None], 0)
    for epi in (False, True):
        got = C0.copy()
        if beta == 0.0:
            got[:] = np.nan
        rc = emu.emu_gemm_skinny_m_f32(batch, M, N, K, alpha, at(A, 0), K, 1, 0, at(B, 0), rsb, csb, bsb, beta, at(got, 0), N, 1,
                                       M * N, 3, at(bias, 0) if epi else None, 1, 1 if epi else 0)
        # 2: the cp.async-staged kernel (B with unit column stride, 16-byte aligned rows), 1: the register-only kernel
        assert rc == (1 if b_col else 2), rc      # (the ragged right edge of a padded B arrives zero-filled)
        want = ref_epi if epi else ref
        if alpha == 1.0:
            assert np.array_equal(got, want)
        else:
            assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max()


def test_f64_tensor_core_kernel_batched(emu):
    batch, M, N, K = 3, 70, 130, 270
    rng = np.random.default_rng(6)
    A = rng.standard_normal((batch, M, K)); B = rng.standard_normal((K, N)); C = rng.standard_normal((batch, M, N)); ref = C.copy()
    for i in range(batch):
        O.gemm_strided(M, N, K, 1.0, A[i], K, 1, B, N, 1, 1.0, ref[i], N, 1)
    emu.emu_gemm_dmma_f64(batch, M, N, K, 1.0, at(A, 0), K, 1, M * K, at(B, 0), N, 1, 0, 1.0, at(C, 0), N, 1, M * N, 4)
    assert np.array_equal(C, ref)


@pytest.mark.parametrize("per_row,act", [(0, 0), (1, 1), (0, 2), (1, 3)])
def test_fused_epilogue_applies_once_after_the_last_k_block(emu, per_row, act):
    M, N, K = 33, 20, 700                    # two kc blocks: the epilogue must run on the second only
    a = O.fill_uniform_f32(M * K, 8, -0.1, 0.1).reshape(M, K); b = O.fill_uniform_f32(K * N, 9, -0.1, 0.1).reshape(K, N)
    bias = O.fill_uniform_f32(M if per_row else N, 10, -1, 1)
    c = np.zeros((M, N), np.float32); ref = c.copy()
    O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, ref, N, 1)
    run(emu, "f32", M, N, K, 1.0, a, 0, K, 1, b, 0, N, 1, 0.0, c, 0, N, 1, epi=(bias, per_row, act))
    v = ref + (bias[:, None] if per_row else bias[None, :])
    exp = {0: v, 1: np.maximum(v, 0), 2: np.tanh(v), 3: 1 / (1 + np.exp(-v))}[act]
    assert np.allclose(c, exp, rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("batch,M,N,K,grid", [(1, 20, 30, 40, 0), (7, 64, 64, 64, 3), (5, 130, 129, 20, 4), (3, 17, 9, 600, 0)])
@pytest.mark.parametrize("shared_b", [False, True])
def test_batched_launch_equals_a_loop_of_gemm_strided(emu, batch, M, N, K, grid, shared_b):
    A = O.fill_uniform_f32(batch * M * K, 11, -1, 1); C = O.fill_uniform_f32(batch * M * N, 13, -1, 1)
    B = O.fill_uniform_f32((1 if shared_b else batch) * K * N, 12, -1, 1)
    bsB = 0 if shared_b else K * N
    ref = C.copy()
    O.gemm_strided_batched(batch, M, N, K, 1.0, A, K, 1, M * K, B, N, 1, bsB, 1.0, ref, N, 1, M * N)
    tiles = emu.emu_gemm_simt_batched_f32(batch, M, N, K, 1.0, at(A, 0), K, 1, M * K, at(B, 0), N, 1, bsB, 1.0, at(C, 0), N, 1,
                                          M * N, grid)
    assert tiles == -(-M // 128) * -(-N // 128)
    assert np.array_equal(C, ref)


def test_batched_f64_and_i64(emu):
    batch, M, N, K = 4, 70, 33, 300
    rng = np.random.default_rng(4)
    A = rng.random(batch * M * K); B = rng.random(batch * K * N); C = rng.random(batch * M * N); ref = C.copy()
    for b in range(batch):
        r = ref[b * M * N:(b + 1) * M * N]
        O.gemm_strided(M, N, K, 1.0, A[b * M * K:], K, 1, B[b * K * N:], N, 1, 1.0, r, N, 1)
    emu.emu_gemm_simt_batched_f64(batch, M, N, K, 1.0, at(A, 0), K, 1, M * K, at(B, 0), N, 1, K * N, 1.0, at(C, 0), N, 1, M * N, 3)
    assert np.array_equal(C, ref)
    Ai = rng.integers(-2**62, 2**62, size=batch * M * K, dtype=np.int64); Bi = rng.integers(-2**62, 2**62, size=K * N, dtype=np.int64)
    Ci = np.zeros(batch * M * N, np.int64); refi = Ci.copy()
    for b in range(batch):
        O.gemm_strided(M, N, K, 1, Ai[b * M * K:], K, 1, Bi, N, 1, 0, refi[b * M * N:(b + 1) * M * N], N, 1)
    emu.emu_gemm_simt_batched_i64(batch, M, N, K, 1, at(Ai, 0), K, 1, M * K, at(Bi, 0), N, 1, 0, 0, at(Ci, 0), N, 1, M * N, 0)
    assert np.array_equal(Ci, refi)


# ---- property-based: random shapes, strides and scalars against the oracle ------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(M=st.integers(1, 140), N=st.integers(1, 140), K=st.integers(1, 560), la=st.sampled_from(LAYOUTS),
       lb=st.sampled_from(LAYOUTS), lc=st.sampled_from(LAYOUTS), beta=st.sampled_from([0.0, 1.0, -0.5]),
       grid=st.integers(0, 4), seed=st.integers(0, 2**31))
def test_property_random_problems_bit_exact(emu, M, N, K, la, lb, lc, beta, grid, seed):
    a = O.fill_uniform_f32(M * K, seed, -1, 1).reshape(M, K); b = O.fill_uniform_f32(K * N, seed + 1, -1, 1).reshape(K, N)
    c0 = O.fill_uniform_f32(M * N, seed + 2, -1, 1).reshape(M, N)
    A, oa, rsa, csa = embed(a, la); B, ob, rsb, csb = embed(b, lb); C, oc, rsc, csc = embed(c0, lc)
    Cref = C.copy()
    O.gemm_strided(M, N, K, 1.0, A[oa:], rsa, csa, B[ob:], rsb, csb, beta, Cref[oc:], rsc, csc)
    run(emu, "f32", M, N, K, 1.0, A, oa, rsa, csa, B, ob, rsb, csb, beta, C, oc, rsc, csc, grid=grid)
    assert np.array_equal(C, Cref)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("NV", [1, 2, 3, 4])
@pytest.mark.parametrize("M,K,grid", [(1, 4, 1), (37, 1000, 2), (300, 4096, 3)])
def test_skinny_gemm_warp_kernels(emu, variant, NV, M, K, grid):
    """N <= 4: one warp per row, shuffle reduction (a different summation tree from the reference's
    k-sequential chain, hence a tolerance; PATH_SIMT never takes these kernels)."""
    a = O.fill_uniform_f32(M * K, 3, -1, 1).reshape(M, K); b = O.fill_uniform_f32(K * NV, 4, -1, 1).reshape(K, NV)
    c = O.fill_uniform_f32(M * NV, 5, -1, 1).reshape(M, NV); ref = c.copy()
    O.gemm_strided(M, NV, K, 0.5, a, K, 1, b, NV, 1, -1.25, ref, NV, 1)
    assert emu.emu_gemv(variant, NV, M, K, 0.5, at(a, 0), K, 1, at(b, 0), NV, 1, -1.25, at(c, 0), NV, 1, grid) == 0
    assert np.abs(c - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()) * np.sqrt(K)
    if variant == 0:      # the scalar variant also takes arbitrary strides: A transposed, C column-major
        at_ = np.ascontiguousarray(a.T); c2 = np.full((NV, M), np.nan, np.float32); ref2 = np.zeros((M, NV), np.float32)
        O.gemm_strided(M, NV, K, 1.0, a, K, 1, b, NV, 1, 0.0, ref2, NV, 1)
        assert emu.emu_gemv(0, NV, M, K, 1.0, at(at_, 0), 1, M, at(b, 0), NV, 1, 0.0, at(c2, 0), 1, M, grid) == 0
        assert np.abs(c2.T - ref2).max() <= 2e-6 * max(1.0, np.abs(ref2).max()) * np.sqrt(K)
