"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on identical inputs.  Bars:
  * integer / SIMT paths: BIT-EXACT with the oracle (same order of operations);
  * fp32 tensor-core, fp32-faithful modes (F16X3 = default, TF32X3): max |ours-ref|/|ref| < 1e-4 on U(0,1) inputs
    (BASELINE.json gate), normwise < 2e-6 and mean_relative_error <= 1e-5 (the reference's
    own gate, gemm_bench_float32.nim:365-367) on U(-0.1,0.1);
  * fp32 tensor-core, 1xTF32 (opt-in fast mode): normwise < 2e-3;
  * bf16: max-elementwise 2^-7 on U(0,1) vs the bf16 oracle (result rounding to bf16 +
    tensor-core accumulation order), normwise < 4e-3.
"""
import numpy as np
import pytest

import oracle as O
from backend import EMU, dev, emu_budget, needs_gpu, sync
from util import LAYOUTS, bf16_bits_to_f32, embed, extract, f32_to_bf16_bits, golden_cases

pytestmark = pytest.mark.gpu

if not EMU:
    import torch
import laser_b200 as L  # noqa: E402

NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}
F32_PATHS = [L.PATH_SIMT, L.PATH_TF32X1, L.PATH_TF32X3, L.PATH_F16X3]
FAITHFUL = (L.PATH_TF32X3, L.PATH_F16X3)   # fp32-faithful tensor-core modes (same gates)


def dptr(t, off, name):
    return L.DevPtr(t.data_ptr() + off * t.element_size(), name)


def run_dev(name, M, N, K, alpha, A, la, B, lb, beta, C0, lc, path=L.PATH_AUTO):
    """Embed logical A, B, C0 in strided buffers, run on the GPU through the _dev C entry,
    return (logical C, whole C buffer after, whole C buffer before, view index mask)."""
    ba, oa, rsa, csa = embed(A, la); bb, ob, rsb, csb = embed(B, lb); bc, oc, rsc, csc = embed(C0, lc)
    view = lambda b: b.view(np.int16) if name == "bf16" else b  # torch has no uint16 math; bits only
    ta, tb, tc = dev(view(ba)), dev(view(bb)), dev(view(bc))
    emu_budget(float(M) * N * K)
    L.gemm_strided(M, N, K, alpha, dptr(ta, oa, name), rsa, csa, dptr(tb, ob, name), rsb, csb, beta,
                   dptr(tc, oc, name), rsc, csc, path=path)
    sync()
    after = tc.cpu().numpy().view(bc.dtype)
    return extract(after, oc, rsc, csc, M, N), after, bc, (oc, rsc, csc)


# --------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_golden_all_dtypes_and_paths(case):
    M, N, K = case["M"], case["N"], case["K"]
    for name in ("f32", "f64", "i32", "i64"):
        a = np.array(case["a"], NP[name]); b = np.array(case["b"], NP[name])
        paths = F32_PATHS + [L.PATH_AUTO] if name == "f32" else [L.PATH_AUTO]
        for path in paths:
            got, *_ = run_dev(name, M, N, K, 1, a, "row", b, "row", 0, np.full((M, N), 99, NP[name]), "row", path)
            assert np.array_equal(got, np.array(case["c"], NP[name])), (name, path)
    a = f32_to_bf16_bits(np.array(case["a"], np.float32)); b = f32_to_bf16_bits(np.array(case["b"], np.float32))
    got, *_ = run_dev("bf16", M, N, K, 1.0, a, "row", b, "row", 0.0, np.zeros((M, N), np.uint16), "row")
    assert np.array_equal(bf16_bits_to_f32(got), np.array(case["c"], np.float32))


@pytest.mark.parametrize("case", golden_cases()[:4], ids=lambda c: c["src"])
def test_golden_host_pointer_entry(case):
    """the drop-in signature itself: numpy (host) buffers in, result in host C."""
    M, N, K = case["M"], case["N"], case["K"]
    for name in ("f32", "f64", "i32", "i64"):
        a = np.array(case["a"], NP[name]); b = np.array(case["b"], NP[name])
        c = np.full((M, N), 99, NP[name])
        L.gemm_strided(M, N, K, 1, a, K, 1, b, N, 1, 0, c, N, 1)
        assert np.array_equal(c, np.array(case["c"], NP[name])), name


# ------------------------------------------------------------------------------ exact path
SHAPES = [(1, 1, 1), (14, 32, 512), (15, 33, 513), (128, 128, 128), (200, 70, 1100), (193, 257, 31),
          (129, 127, 1025), (64, 300, 2049)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ab", [(1.0, 0.0), (0.5, -1.25), (1.0, 1.0)])
def test_simt_bit_exact_f32(shape, ab):
    M, N, K = shape
    alpha, beta = ab
    A = O.fill_uniform_f32(M * K, 11, -0.1, 0.1).reshape(M, K); B = O.fill_uniform_f32(K * N, 12, -0.1, 0.1).reshape(K, N)
    C0 = O.fill_uniform_f32(M * N, 13, -1, 1).reshape(M, N)
    want = C0.copy(); O.gemm_strided(M, N, K, alpha, A, K, 1, B, N, 1, beta, want, N, 1)
    got, *_ = run_dev("f32", M, N, K, alpha, A, "row", B, "row", beta, C0, "row", L.PATH_SIMT)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("name", ["f64", "i32", "i64"])
def test_simt_bit_exact_other_dtypes(name):
    M, N, K = 77, 45, 700
    rng = np.random.default_rng(3)
    if name == "f64":
        A = rng.uniform(-1, 1, (M, K)); B = rng.uniform(-1, 1, (K, N)); C0 = rng.uniform(-1, 1, (M, N)); al, be = 0.5, -1.25
    else:
        A = rng.integers(-2**20, 2**20, (M, K)).astype(NP[name]); B = rng.integers(-2**20, 2**20, (K, N)).astype(NP[name])
        C0 = rng.integers(-100, 100, (M, N)).astype(NP[name]); al, be = 3, -2
    want = C0.copy(); O.gemm_strided(M, N, K, al, A, K, 1, B, N, 1, be, want, N, 1)
    got, *_ = run_dev(name, M, N, K, al, A, "col", B, "colslice", be, C0, "padded")
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,ab,lay", [((1200, 1100, 700), (1.0, 0.0), ("row", "row", "row")),
                                          ((1153, 1290, 515), (0.5, -1.25), ("col", "colslice", "padded")),
                                          ((2048, 1024, 256), (1.0, 1.0), ("row", "col", "col"))])
def test_f64_dmma_bit_exact(shape, ab, lay):
    """float64 problems whose 128 x 128 tiles fill half of the SMs run on the fp64 tensor cores (mma.sync.m8n8k4.f64,
    csrc/gemm_dmma.cuh).  One DMMA performs, per output element, four steps of the reference's k-sequential FMA chain, so the
    result must equal the oracle's BIT FOR BIT (kc = 256 block boundaries, ragged tiles, strided operands, alpha / beta)."""
    if EMU:
        pytest.skip("the emulated library models the kernel in tests/test_emulated_simt.py")
    M, N, K = shape
    alpha, beta = ab
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, (M, K)); B = rng.uniform(-1, 1, (K, N)); C0 = rng.uniform(-1, 1, (M, N))
    want = C0.copy(); O.gemm_strided(M, N, K, alpha, A, K, 1, B, N, 1, beta, want, N, 1)
    n0 = L.lib().laser_b200_debug_f64_dmma_launches()
    got, after, before, (oc, rsc, csc) = run_dev("f64", M, N, K, alpha, A, lay[0], B, lay[1], beta, C0, lay[2])
    assert L.lib().laser_b200_debug_f64_dmma_launches() == n0 + 1, "the problem was expected to take the DMMA kernel"
    if alpha == 1.0:
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    else:      # alpha != 1: the oracle's compiler may contract C += alpha * AB into one fma (1 ulp), as for f32
        assert np.abs(got - want).max() <= 4e-16 * np.abs(want).max()
    mask = np.ones(after.shape, bool).reshape(-1)
    idx = oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc
    mask[idx.reshape(-1)] = False
    assert np.array_equal(after.reshape(-1)[mask], before.reshape(-1)[mask])      # nothing outside the view changed


@pytest.mark.parametrize("la", LAYOUTS)
@pytest.mark.parametrize("lb", ["row", "col", "colslice", "negrow"])
@pytest.mark.parametrize("lc", ["row", "col", "padded", "negcol"])
@pytest.mark.parametrize("path", F32_PATHS)
def test_every_stride_class(la, lb, lc, path):
    """arbitrary strides on A, B and C for all three fp32 kernel families; elements of the C
    buffer outside the view must not change."""
    M, N, K = 150, 270, 200
    A = O.fill_uniform_f32(M * K, 21, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 22, 0, 1).reshape(K, N)
    C0 = O.fill_uniform_f32(M * N, 23, 0, 1).reshape(M, N)
    want = C0.copy(); O.gemm_strided(M, N, K, 0.75, A, K, 1, B, N, 1, 0.5, want, N, 1)
    got, after, before, (oc, rsc, csc) = run_dev("f32", M, N, K, 0.75, A, la, B, lb, 0.5, C0, lc, path)
    if path == L.PATH_SIMT:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    else:
        tol = 1e-4 if path in FAITHFUL else 5e-3
        assert O.max_relative_error(got, want) < tol, O.max_relative_error(got, want)
    mask = np.ones(after.size, bool); mask[(oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).ravel()] = False
    assert np.array_equal(after[mask], before[mask])


# ------------------------------------------------------------------- tensor-core tolerances
TC_SHAPES = [(128, 256, 32), (256, 512, 4096), (300, 500, 1000), (1000, 777, 513), (4096, 128, 64), (129, 257, 8200)]


@pytest.mark.parametrize("path", FAITHFUL)
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_faithful_modes_meet_fp32_gates(shape, path):
    M, N, K = shape
    for seed, lo, hi in ((42, 0.0, 1.0), (42, -0.1, 0.1)):
        A = O.fill_uniform_f32(M * K, seed, lo, hi).reshape(M, K); B = O.fill_uniform_f32(K * N, seed + 1, lo, hi).reshape(K, N)
        want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
        got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.full((M, N), np.nan, np.float32), "row", path)
        assert O.normwise_relative_error(got, want) < 2e-6
        if lo >= 0:
            assert O.max_relative_error(got, want) < 1e-4     # BASELINE.json gate (P inputs)
        assert O.mean_relative_error(got, want) <= 1e-5       # reference's own gate (S inputs too)


@pytest.mark.parametrize("shape", TC_SHAPES[:4])
def test_tf32x1_fast_mode_tolerance(shape):
    M, N, K = shape
    A = O.fill_uniform_f32(M * K, 7, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 8, 0, 1).reshape(K, N)
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
    got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row", L.PATH_TF32X1)
    assert O.normwise_relative_error(got, want) < 2e-3


@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 500, 1000), (1000, 777, 513)])
@pytest.mark.parametrize("lay", [("row", "row"), ("col", "row"), ("row", "col"), ("col", "col"), ("padded", "colslice")])
def test_bf16(shape, lay):
    M, N, K = shape
    A = f32_to_bf16_bits(O.fill_uniform_f32(M * K, 31, 0, 1)).reshape(M, K); B = f32_to_bf16_bits(O.fill_uniform_f32(K * N, 32, 0, 1)).reshape(K, N)
    C0 = f32_to_bf16_bits(O.fill_uniform_f32(M * N, 33, 0, 1)).reshape(M, N)
    for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
        want = C0.copy(); O.gemm_strided(M, N, K, alpha, A, K, 1, B, N, 1, beta, want, N, 1, bf16=True)
        got, *_ = run_dev("bf16", M, N, K, alpha, A, lay[0], B, lay[1], beta, C0, "row")
        g, w = bf16_bits_to_f32(got), bf16_bits_to_f32(want)
        assert O.max_relative_error(g, w) <= 2.0 ** -7
        assert O.normwise_relative_error(g, w) < 4e-3


@pytest.mark.parametrize("shape", [(256, 256, 8192), (300, 520, 4096), (128, 2048, 16384), (1000, 1000, 3000)])
@pytest.mark.parametrize("path", FAITHFUL)
def test_split_k_small_mn_long_k(shape, path):
    """few output tiles + long K: the K range is split over idle SMs, partial planes are reduced in
    a fixed order by a second kernel (deterministic), alpha/beta/strided C applied there."""
    M, N, K = shape
    A = O.fill_uniform_f32(M * K, 5, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 6, 0, 1).reshape(K, N)
    C0 = O.fill_uniform_f32(M * N, 7, 0, 1).reshape(M, N)
    want = C0.copy(); O.gemm_strided(M, N, K, 0.5, A, K, 1, B, N, 1, -1.25, want, N, 1)
    n0 = L.launch_count()
    got, *_ = run_dev("f32", M, N, K, 0.5, A, "row", B, "row", -1.25, C0, "padded", path)
    launches = L.launch_count() - n0
    assert O.max_relative_error(got, want) < 1e-4 and O.normwise_relative_error(got, want) < 2e-6
    again, *_ = run_dev("f32", M, N, K, 0.5, A, "row", B, "row", -1.25, C0, "padded", path)
    assert np.array_equal(got, again)                      # deterministic reduction order
    if shape != (1000, 1000, 3000):
        # TF32X3: split A, split B, GEMM, reduce.  F16X3: fused scale+split of A, abs-max + split of the MN-major B, GEMM, reduce
        assert launches == (5 if path == L.PATH_F16X3 else 4)


def test_beta_zero_never_reads_c():
    M, N, K = 130, 260, 600
    A = O.fill_uniform_f32(M * K, 4, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 5, 0, 1).reshape(K, N)
    for path in F32_PATHS:
        clean, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row", path)
        for fill in (np.nan, np.inf):
            got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.full((M, N), fill, np.float32), "row", path)
            assert np.array_equal(got, clean), path


def test_degenerate_extents_and_errors():
    c = dev(np.arange(6, dtype=np.float32)); a = dev(np.zeros(8, np.float32))
    pa, pc = dptr(a, 0, "f32"), dptr(c, 0, "f32")
    L.gemm_strided(2, 3, 0, 1.0, pa, 0, 1, pa, 3, 1, 0.5, pc, 3, 1)     # K == 0: C untouched (gemm.nim:150)
    L.gemm_strided(0, 3, 2, 1.0, pa, 2, 1, pa, 3, 1, 0.5, pc, 3, 1)
    sync()
    assert np.array_equal(c.cpu().numpy(), np.arange(6, dtype=np.float32))
    with pytest.raises(L.LaserB200Error):
        L.gemm_strided(-1, 3, 2, 1.0, pa, 2, 1, pa, 3, 1, 0.5, pc, 3, 1)


# ------------------------------------------------------------------- host-pointer drop-in
@pytest.mark.parametrize("la,lb,lc", [("row", "row", "row"), ("col", "colslice", "padded"), ("negrow", "negcol", "negrow")])
def test_host_pointer_entry_strided(la, lb, lc):
    M, N, K = 300, 280, 520
    A = O.fill_uniform_f32(M * K, 51, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 52, 0, 1).reshape(K, N)
    C0 = O.fill_uniform_f32(M * N, 53, 0, 1).reshape(M, N)
    for alpha, beta in ((1.0, 0.0), (0.5, -1.25)):
        want = C0.copy(); O.gemm_strided(M, N, K, alpha, A, K, 1, B, N, 1, beta, want, N, 1)
        ba, oa, rsa, csa = embed(A, la); bb, ob, rsb, csb = embed(B, lb); bc, oc, rsc, csc = embed(C0, lc)
        before = bc.copy()
        L.gemm_strided(M, N, K, alpha, ba[oa:], rsa, csa, bb[ob:], rsb, csb, beta, bc[oc:], rsc, csc)
        assert L.last_path() == L.PATH_F16X3     # the default fp32-faithful mode
        got = extract(bc, oc, rsc, csc, M, N)
        assert O.max_relative_error(got, want) < 1e-4
        mask = np.ones(bc.size, bool); mask[(oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).ravel()] = False
        assert np.array_equal(bc[mask], before[mask])


@pytest.mark.parametrize("la,lc,M", [("row", "row", 2500), ("padded", "padded", 2500), ("col", "row", 2500),
                                     ("row", "row", 2100), ("row", "row", 2049)])
def test_host_pointer_entry_pipelined(la, lc, M):
    """M >= 2048 with separable row panels takes the 3-stream pipelined host path (H2D of panel
    p+1 | split+GEMM of panel p | D2H of panel p-1); 'col' A is not separable -> plain path.
    2100 / 2049: the last panel is shorter than one 128-row tile."""
    N, K = 520, 300
    A = O.fill_uniform_f32(M * K, 81, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 82, 0, 1).reshape(K, N)
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 0.5, A, K, 1, B, N, 1, 0.0, want, N, 1)
    ba, oa, rsa, csa = embed(A, la); bc, oc, rsc, csc = embed(np.full((M, N), np.nan, np.float32), lc)
    before = bc.copy()
    L.gemm_strided(M, N, K, 0.5, ba[oa:], rsa, csa, B, N, 1, 0.0, bc[oc:], rsc, csc)
    got = extract(bc, oc, rsc, csc, M, N)
    assert O.max_relative_error(got, want) < 1e-4
    mask = np.ones(bc.size, bool); mask[(oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).ravel()] = False
    assert np.array_equal(bc[mask], before[mask], equal_nan=True)


def test_auto_path_selection():
    a = dev(np.ones(128 * 128, np.float32)); c = dev(np.zeros(128 * 128, np.float32))
    L.gemm_strided(128, 128, 128, 1.0, dptr(a, 0, "f32"), 128, 1, dptr(a, 0, "f32"), 128, 1, 0.0, dptr(c, 0, "f32"), 128, 1)
    assert L.last_path() == L.PATH_SIMT            # M*N*K <= 128^3: exact kernel (gemm.nim:140-141 threshold)
    a = dev(np.ones(256 * 256, np.float32)); c = dev(np.zeros(256 * 256, np.float32))
    L.gemm_strided(256, 256, 256, 1.0, dptr(a, 0, "f32"), 256, 1, dptr(a, 0, "f32"), 256, 1, 0.0, dptr(c, 0, "f32"), 256, 1)
    assert L.last_path() == L.PATH_F16X3
    sync()
    assert np.all(c.cpu().numpy() == 256.0)
    # few output rows, wide N (the im2col convolution's product): exact few-rows kernel, whatever the tensor-core mode
    a = dev(np.ones(20 * 300, np.float32)); b = dev(np.ones(300 * 2048, np.float32)); c = dev(np.zeros(20 * 2048, np.float32))
    L.gemm_strided(20, 2048, 300, 1.0, dptr(a, 0, "f32"), 300, 1, dptr(b, 0, "f32"), 2048, 1, 0.0, dptr(c, 0, "f32"), 2048, 1)
    assert L.last_path() == L.PATH_SIMT
    sync()
    assert np.all(c.cpu().numpy() == 300.0)


@pytest.mark.parametrize("K", [777, 776])      # 776: the 16-byte vectorised variant
def test_skinny_gemv_path(K):
    M, N = 5000, 3
    A = O.fill_uniform_f32(M * K, 61, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 62, 0, 1).reshape(K, N)
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
    got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row")
    assert O.max_relative_error(got, want) < 1e-5


def test_simt_mode_is_exact_for_every_shape_and_both_entries():
    """The SIMT mode is documented as bit-identical to the CPU reference: it must win over the GEMV shortcut (N <= 4,
    tall) of PATH_AUTO, and the host-pointer entry must pick the same kernel family as the device entry (a 2048 x 8 x 8
    problem is below the 128^3 switch: exact kernel, not the pipelined tensor-core path)."""
    old = L.get_f32_mode()
    try:
        L.set_f32_mode(L.PATH_SIMT)
        M, N, K = 1500, 3, 700
        A = O.fill_uniform_f32(M * K, 91, -1, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 92, -1, 1).reshape(K, N)
        want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
        got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row")
        assert L.last_path() == L.PATH_SIMT and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finally:
        L.set_f32_mode(old)
    M, N, K = 2048, 8, 8
    A = O.fill_uniform_f32(M * K, 93, -1, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 94, -1, 1).reshape(K, N)
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
    host = np.full((M, N), np.nan, np.float32)
    L.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, host, N, 1)            # host-pointer entry
    assert L.last_path() == L.PATH_SIMT and np.array_equal(host.view(np.uint32), want.view(np.uint32))
    got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row")
    assert L.last_path() == L.PATH_SIMT and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("path", [L.PATH_F16X3, L.PATH_TF32X3])
def test_infinite_entries_do_not_poison_other_outputs(path):
    """x = +-inf in an operand: the split modes compute x - hi = inf - inf for the low piece unless it is forced to 0; a
    NaN there would turn a whole row / column of C into NaN, where the reference gives +-inf only in the outputs the
    infinite entry reaches with a non-zero partner (and finite values everywhere else)."""
    M, N, K = 200, 300, 160
    A = O.fill_uniform_f32(M * K, 95, 0.5, 1).reshape(M, K).copy(); B = O.fill_uniform_f32(K * N, 96, 0.5, 1).reshape(K, N).copy()
    A[7, 11] = np.inf; B[13, 21] = -np.inf
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
    got, *_ = run_dev("f32", M, N, K, 1.0, A, "row", B, "row", 0.0, np.zeros((M, N), np.float32), "row", path)
    fin = np.isfinite(want)
    assert fin.sum() == (M - 1) * (N - 1)
    assert np.isfinite(got[fin]).all()                          # rows / columns the infinities do not touch stay finite
    assert O.max_relative_error(got[fin], want[fin]) < 1e-4
    # all operands positive, so every output the infinite entries reach is that infinity in the reference (and +inf - inf
    # = NaN at their crossing); here they must be non-finite as well -- never a finite number
    assert not np.isfinite(got[~fin]).any()


# --------------------------------------------------------------------------- Tensor contract
def test_tensor_contract_and_matmul():
    M, N, K = 200, 260, 300
    A = O.fill_uniform_f32(M * K, 71, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 72, 0, 1).reshape(K, N)
    tA, tBt = L.toTensor(A), L.toTensor(np.ascontiguousarray(B.T))
    assert tA.rank == 2 and tA.size == M * K and tA.is_C_contiguous() and tA.strides == [K, 1]
    tB = tBt.transpose()                               # a strided view sharing storage
    assert tB.shape == [K, N] and tB.strides == [1, K] and not tB.is_C_contiguous()
    assert np.array_equal(tB.to_numpy(), B)
    tC = L.matmul(tA, tB)
    want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
    assert O.max_relative_error(tC.to_numpy(), want) < 1e-4
    z = L.newTensor([3, 5])
    assert np.all(z.to_numpy() == 0)                   # newTensor zero-initialises (initialization.nim:156-170)
    sub = tA.slice2d(slice(10, 110), slice(0, K, 2))   # A[10:110, ::2]
    assert np.array_equal(sub.to_numpy(), A[10:110, ::2])
    tC2 = L.matmul(sub, L.toTensor(B[::2].copy()), path=L.PATH_SIMT)
    want2 = np.zeros((100, N), np.float32); Asub = np.ascontiguousarray(A[10:110, ::2]); Bsub = B[::2].copy()
    O.gemm_strided(100, N, Asub.shape[1], 1.0, Asub, Asub.shape[1], 1, Bsub, N, 1, 0.0, want2, N, 1)
    assert np.array_equal(tC2.to_numpy(), want2)
    if not EMU:
        tt = L.Tensor.from_torch(torch.arange(12, dtype=torch.float32, device="cuda").reshape(3, 4)[:, 1:])
        assert tt.shape == [3, 3] and tt.strides == [4, 1] and tt.offset == 1
        assert np.array_equal(tt.to_numpy(), np.arange(12, dtype=np.float32).reshape(3, 4)[:, 1:])


@needs_gpu
def test_device_fill_matches_oracle_bit_for_bit():
    n = 100003
    t = torch.empty(n, dtype=torch.float32, device="cuda")
    L.fill_uniform_f32(t, n, 42, -0.1, 0.1)
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy(), O.fill_uniform_f32(n, 42, -0.1, 0.1))


# --------------------------------------------------------- BASELINE.json full-size properties
def _rows_check(M, N, K, tA, tB, tC, rows, tol):
    A_rows = tA[rows].cpu().numpy(); Bh = tB.cpu().numpy()
    want = np.zeros((len(rows), N), np.float32)
    O.gemm_strided(len(rows), N, K, 1.0, A_rows, K, 1, Bh, N, 1, 0.0, want, N, 1)
    got = tC[rows].cpu().numpy()
    assert O.max_relative_error(got, want) < tol, O.max_relative_error(got, want)


@needs_gpu
@pytest.mark.parametrize("n", [4096, 8192])
def test_full_size_sgemm_sampled_rows(n):
    """configs[1] (4096^3) and the metric shape (8192^3): device-generated U(0,1) inputs; 48
    sampled rows of C are checked against the oracle, and the SIMT kernel cross-checks a
    256-row panel bit-for-bit against the oracle order."""
    M = N = K = n
    tA = torch.empty(M * K, dtype=torch.float32, device="cuda"); tB = torch.empty(K * N, dtype=torch.float32, device="cuda")
    L.fill_uniform_f32(tA, M * K, 42, 0, 1); L.fill_uniform_f32(tB, K * N, 43, 0, 1)
    tC = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, tC, N, 1)
    assert L.last_path() == L.PATH_F16X3
    torch.cuda.synchronize()
    assert not torch.isnan(tC).any()
    rows = np.unique(np.random.default_rng(0).integers(0, M, 48))
    _rows_check(M, N, K, tA.view(M, K), tB.view(K, N), tC, rows, 1e-4)
    # linearity: C(2A, B) == 2*C(A, B) exactly (scaling by 2 is exact in every pass)
    tC2 = torch.empty_like(tC)
    L.gemm_strided(M, N, K, 1.0, tA * 2, K, 1, tB, N, 1, 0.0, tC2, N, 1)
    torch.cuda.synchronize()
    assert torch.equal(tC2, tC * 2)
    if n == 4096:
        panel = torch.zeros((256, N), dtype=torch.float32, device="cuda")
        L.gemm_strided(256, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, panel, N, 1, path=L.PATH_SIMT)
        torch.cuda.synchronize()
        want = np.zeros((256, N), np.float32)
        O.gemm_strided(256, N, K, 1.0, tA.view(M, K)[:256].cpu().numpy(), K, 1, tB.view(K, N).cpu().numpy(), N, 1, 0.0, want, N, 1)
        assert np.array_equal(panel.cpu().numpy().view(np.uint32), want.view(np.uint32))


def _gates_on_device(got, ref, positive):
    """the three fp32 gates on whole device matrices (float64 accumulation on the GPU): max-elementwise relative error
    (BASELINE.json, positive inputs only), normwise, and the reference's own mean_relative_error
    (laser/private/error_functions.nim:6-26: |y - y_true| / max(|y|, |y_true|), 0 when both are 0)"""
    g, r = got.double(), ref.double()
    d = (g - r).abs()
    normwise = (torch.linalg.norm(g - r) / torch.linalg.norm(r)).item()
    den = torch.maximum(g.abs(), r.abs())
    mre = torch.where(den > 0, d / den, torch.zeros_like(d)).mean().item()
    max_rel = (d / r.abs()).max().item() if positive else None
    return max_rel, normwise, mre


@needs_gpu
@pytest.mark.parametrize("shape", [(8192, 8192, 8192), (4096, 4096, 16384)], ids=["8192^3", "4096x4096x16384"])
@pytest.mark.parametrize("dist", ["P_U(0,1)", "S_U(-0.1,0.1)"])
def test_default_mode_meets_all_gates_at_the_metric_shape(shape, dist):
    """Where the metric is quoted (8192^3) and at twice that K: the DEFAULT fp32 mode against the reference's order of
    operations on the WHOLE matrix, both input distributions of the reference's benches (P = U(0,1),
    gemm_bench_float64.nim:198-199; S = U(-0.1,0.1), gemm_bench_float32.nim:343-344), all three gates: max-elementwise
    < 1e-4 (P), normwise < 2e-6, mean_relative_error <= 1e-5 (gemm_bench_float32.nim:365-367).  The reference result is
    the exact SIMT kernel's (same blocked-K fmaf chain as the CPU reference); 512 sampled rows of it are pinned bit for
    bit to the CPU restatement of the reference at this very size."""
    M, N, K = shape
    lo, hi = (0.0, 1.0) if dist.startswith("P") else (-0.1, 0.1)
    tA = torch.empty(M * K, dtype=torch.float32, device="cuda"); tB = torch.empty(K * N, dtype=torch.float32, device="cuda")
    L.fill_uniform_f32(tA, M * K, 42, lo, hi); L.fill_uniform_f32(tB, K * N, 43, lo, hi)
    got = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, got, N, 1)
    assert L.last_path() == L.PATH_F16X3
    ref = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, ref, N, 1, path=L.PATH_SIMT)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any() and not torch.isnan(ref).any()
    max_rel, normwise, mre = _gates_on_device(got, ref, lo >= 0)
    assert normwise < 2e-6, normwise
    assert mre <= 1e-5, mre
    if max_rel is not None:
        assert max_rel < 1e-4, max_rel
    # pin the SIMT result to the CPU restatement of the reference (oracle/laser_cpu_gemm.c, bit-equal to the numerics oracle:
    # tests/test_oracle.py) on 512 sampled rows
    rows = np.unique(np.random.default_rng(7).integers(0, M, 512))
    A_rows = np.ascontiguousarray(tA.view(M, K)[torch.as_tensor(rows, device="cuda")].cpu().numpy()); Bh = tB.view(K, N).cpu().numpy()
    want = np.zeros((len(rows), N), np.float32)
    O.cpu_gemm_strided_f32(len(rows), N, K, 1.0, A_rows.reshape(-1), K, 1, Bh.reshape(-1), N, 1, 0.0, want.reshape(-1), N, 1)
    ref_rows = ref[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    assert np.array_equal(ref_rows.view(np.uint32), want.view(np.uint32))
    got_rows = got[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    assert O.normwise_relative_error(got_rows, want) < 2e-6 and O.mean_relative_error(got_rows, want) <= 1e-5
    if lo >= 0:
        assert O.max_relative_error(got_rows, want) < 1e-4


@needs_gpu
def test_full_size_transposed_a_4096():
    """configs[2]: A given transposed (storage K x M, rowStrideA = 1, colStrideA = M)."""
    M = N = K = 4096
    tAt = torch.empty(K * M, dtype=torch.float32, device="cuda"); tB = torch.empty(K * N, dtype=torch.float32, device="cuda")
    L.fill_uniform_f32(tAt, K * M, 44, 0, 1); L.fill_uniform_f32(tB, K * N, 45, 0, 1)
    for path in (L.PATH_F16X3, L.PATH_TF32X3, L.PATH_TF32X1):
        tC = torch.empty((M, N), dtype=torch.float32, device="cuda")
        L.gemm_strided(M, N, K, 1.0, tAt, 1, M, tB, N, 1, 0.0, tC, N, 1, path=path)
        torch.cuda.synchronize()
        rows = np.unique(np.random.default_rng(1).integers(0, M, 32))
        A_logical = tAt.view(K, M).t()
        _rows_check(M, N, K, A_logical, tB.view(K, N), tC, rows, 1e-4 if path in FAITHFUL else 5e-3)


@needs_gpu
def test_full_size_bf16_8192():
    """configs[3]: bf16 8192^3, sampled rows vs the bf16 oracle."""
    M = N = K = 8192
    tA = (torch.rand(M, K, device="cuda")).to(torch.bfloat16); tB = (torch.rand(K, N, device="cuda")).to(torch.bfloat16)
    tC = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    L.gemm_strided(M, N, K, 1.0, tA, K, 1, tB, N, 1, 0.0, tC, N, 1)
    torch.cuda.synchronize()
    rows = np.unique(np.random.default_rng(2).integers(0, M, 16))
    a = tA[rows].view(torch.int16).cpu().numpy().view(np.uint16); b = tB.view(torch.int16).cpu().numpy().view(np.uint16)
    want = np.zeros((len(rows), N), np.uint16)
    O.gemm_strided(len(rows), N, K, 1.0, a, K, 1, b, N, 1, 0.0, want, N, 1, bf16=True)
    got = tC[rows].view(torch.int16).cpu().numpy().view(np.uint16)
    assert O.max_relative_error(bf16_bits_to_f32(got), bf16_bits_to_f32(want)) <= 2.0 ** -7


@pytest.mark.parametrize("mode,want_path,tol", [("simt", "PATH_SIMT", 0.0), ("tf32x3", "PATH_TF32X3", 1e-4),
                                                 ("tf32x1", "PATH_TF32X1", 5e-3), ("f16x3", "PATH_F16X3", 1e-4), (None, "PATH_F16X3", 1e-4)])
def test_env_selects_f32_mode(mode, want_path, tol):
    """LASER_B200_F32_MODE picks the kernel family of the drop-in (host-pointer) call; 'simt' makes
    it bit-identical to the CPU reference order."""
    import os, subprocess, sys
    code = f"""
import sys, numpy as np
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import laser_b200 as L, oracle as O
M, N, K = 300, 280, 520
A = O.fill_uniform_f32(M * K, 1, 0, 1).reshape(M, K); B = O.fill_uniform_f32(K * N, 2, 0, 1).reshape(K, N)
want = np.zeros((M, N), np.float32); O.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, want, N, 1)
C = np.zeros((M, N), np.float32); L.gemm_strided(M, N, K, 1.0, A, K, 1, B, N, 1, 0.0, C, N, 1)
assert L.last_path() == L.{want_path}, L.last_path()
err = O.max_relative_error(C, want)
assert err <= {tol}, err
print("ok", err)
"""
    env = dict(os.environ)
    env.pop("LASER_B200_F32_MODE", None)
    if mode is not None:          # None: the built-in default
        env["LASER_B200_F32_MODE"] = mode
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
