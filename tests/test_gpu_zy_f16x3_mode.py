"""GPU parity tests of the DEFAULT fp32 mode, LASER_B200_PATH_F16X3 (the same file runs against the host-emulated library
in the CPU suite, tests/test_emulated_python_mirror.py).

Every row of A / column of B is multiplied by its own 2^s (s from a device-side abs-max, csrc/f16_scale.cuh) and split
into two FP16 arrays, |x 2^s - h - l| <= 2^-22 |x 2^s|; the kernel runs the three-pass order h*l', l*h', h*h' (fp32
output, kc-blocked fp32 accumulation) and its epilogue undoes the scales.  Bars: those of the fp32-faithful modes of
tests/test_gpu_parity.py (U(0,1) max-elementwise < 1e-4, expected ~1e-6; U(-0.1,0.1) normwise < 2e-6,
mean_relative_error <= 1e-5), |ours-exact| <= (3*2^-22 + 2e-6) * sum_k |a||b|; operands far outside fp16's range; one
scale per row of A / column of B; entries below 2^-17 of their own row's / column's maximum keep absolute (not
relative) precision.
"""
import numpy as np
import pytest

import oracle as O
from backend import EMU, dev, emu_budget, sync
from util import LAYOUTS, embed, extract, golden_cases

pytestmark = pytest.mark.gpu

import laser_b200 as L  # noqa: E402


def dptr(t, off=0):
    return L.DevPtr(t.data_ptr() + off * t.element_size(), "f32")


MODES = [L.PATH_F16X3]
PER_PRODUCT = {L.PATH_F16X3: 3 * 2.0 ** -22}
NORMWISE_S = {L.PATH_F16X3: 2e-6}
LAYOUT_TOL = {L.PATH_F16X3: 3e-6}
mode_ids = lambda p: L.PATH_NAMES[p]


def run(path, M, N, K, alpha, a, la, b, lb, beta, c0, lc):
    ba, oa, rsa, csa = embed(a, la); bb, ob, rsb, csb = embed(b, lb); bc, oc, rsc, csc = embed(c0, lc)
    ta, tb, tc = dev(ba), dev(bb), dev(bc)
    emu_budget(3.0 * M * N * K)
    L.gemm_strided(M, N, K, alpha, dptr(ta, oa), rsa, csa, dptr(tb, ob), rsb, csb, beta, dptr(tc, oc), rsc, csc, path=path)
    sync()
    assert L.last_path() == path
    after = tc.cpu().numpy()
    return extract(after, oc, rsc, csc, M, N), after, bc, (oc, rsc, csc)


def bound(path, alpha, a, b, exact):
    return abs(alpha) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)) * (PER_PRODUCT[path] + 2e-6) + np.abs(exact) * 2e-6 + 1e-30


@pytest.mark.parametrize("path", MODES, ids=mode_ids)
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["src"])
def test_golden_vectors_exact(case, path):
    """small integers are exactly representable in one piece (times a power of two): the reference's known answers
    come out exactly"""
    M, N, K = case["M"], case["N"], case["K"]
    a = np.array(case["a"], np.float32); b = np.array(case["b"], np.float32)
    got, *_ = run(path, M, N, K, 1.0, a, "row", b, "row", 0.0, np.full((M, N), 99, np.float32), "row")
    assert np.array_equal(got, np.array(case["c"], np.float32))


SHAPES = [(129, 257, 100), (256, 256, 128), (257, 260, 129), (300, 9, 333), (128, 512, 1100), (513, 300, 2049)]


@pytest.mark.parametrize("path", MODES, ids=mode_ids)
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ab", [(1.0, 0.0), (0.5, -1.25)])
def test_error_bound_and_gates(shape, ab, path):
    M, N, K = shape
    alpha, beta = ab
    for dist, (lo, hi) in (("P", (0.0, 1.0)), ("S", (-0.1, 0.1))):
        a = O.fill_uniform_f32(M * K, 51, lo, hi).reshape(M, K); b = O.fill_uniform_f32(K * N, 52, lo, hi).reshape(K, N)
        c0 = O.fill_uniform_f32(M * N, 53, lo, hi).reshape(M, N)
        start = c0 if beta else np.full((M, N), np.nan, np.float32)          # beta == 0 must not read C
        got, *_ = run(path, M, N, K, alpha, a, "row", b, "row", beta, start, "row")
        exact = alpha * (a.astype(np.float64) @ b.astype(np.float64)) + beta * c0.astype(np.float64)
        err = np.abs(got - exact)
        assert (err <= bound(path, alpha, a, b, exact)).all(), (dist, float((err / bound(path, alpha, a, b, exact)).max()))
        if dist == "P" and beta == 0.0:
            ref = np.zeros((M, N), np.float32)
            O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, 0.0, ref, N, 1)
            assert (np.abs(got - ref) / np.abs(ref)).max() < 1e-4
        if dist == "S":
            assert np.linalg.norm(got - exact) / np.linalg.norm(exact) < NORMWISE_S[path]
            if path == L.PATH_F16X3 and beta == 0.0:      # the reference's own statistical gate (gemm_bench_float32.nim:365-367)
                ref = np.zeros((M, N), np.float32)
                O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, 0.0, ref, N, 1)
                assert np.mean(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)) <= 1e-5


@pytest.mark.parametrize("path", MODES, ids=mode_ids)
@pytest.mark.parametrize("which", ["A", "B", "C"])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_every_operand_class(which, layout, path):
    """K-major operands go through the fused scale + split kernel, MN-major ones through abs-max + split, general strides
    through the gather to fp32 first; C of any strides; nothing outside the C view is written"""
    M, N, K = 150, 140, 100
    a = O.fill_uniform_f32(M * K, 54, 0, 1).reshape(M, K); b = O.fill_uniform_f32(K * N, 55, 0, 1).reshape(K, N)
    c0 = O.fill_uniform_f32(M * N, 56, 0, 1).reshape(M, N)
    la, lb, lc = (layout, "row", "row") if which == "A" else (("row", layout, "row") if which == "B" else ("col", "col", layout))
    got, after, before, (oc, rsc, csc) = run(path, M, N, K, 1.0, a, la, b, lb, 2.0, c0, lc)
    ref = c0.copy()
    O.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 2.0, ref, N, 1)
    assert np.abs(got - ref).max() <= LAYOUT_TOL[path] * np.abs(ref).max()
    idx = (oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc).reshape(-1)
    mask = np.ones(after.size, bool); mask[idx] = False
    assert np.array_equal(after.reshape(-1)[mask], before.reshape(-1)[mask])


@pytest.mark.parametrize("path", MODES, ids=mode_ids)
def test_mode_selection_and_host_pointer_entry(path):
    """set_f32_mode(mode) makes it the AUTO path of device- and host-pointer calls (the pipelined row-panel path
    included: every row panel of A gets its own scale); another mode set before is replaced, and F16X3 is what a fresh
    process starts with (tests/test_gpu_parity.py: test_env_selects_f32_mode)"""
    L.set_f32_mode(L.PATH_TF32X3)
    assert L.get_f32_mode() == L.PATH_TF32X3
    L.set_f32_mode(path)
    try:
        assert L.get_f32_mode() == path
        for (M, N, K) in ((300, 70, 200), (2304, 24, 64)):     # above the 128^3 threshold of the exact kernel
            a = O.fill_uniform_f32(M * K, 57, 0, 1).reshape(M, K); b = O.fill_uniform_f32(K * N, 58, 0, 1).reshape(K, N)
            c0 = O.fill_uniform_f32(M * N, 59, 0, 1).reshape(M, N)
            for alpha, beta in ((1.0, 0.0), (0.5, -1.25)):
                c = c0.copy() if beta else np.full((M, N), np.nan, np.float32)
                L.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, c, N, 1)       # host pointers
                assert L.last_path() == path
                ref = c0.copy()
                O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, ref, N, 1)
                assert np.abs(c - ref).max() <= LAYOUT_TOL[path] * np.abs(ref).max(), (M, N, K, alpha, beta)
    finally:
        L.set_f32_mode(L.PATH_F16X3)
    assert L.get_f32_mode() == L.PATH_F16X3


def test_f16x3_range_handling():
    """fp16 has 5 exponent bits: operands far outside its range in both directions, a zero operand, rows and columns of
    wildly different magnitude, and the documented behaviour for entries far below their own row's maximum"""
    M, N, K = 200, 130, 96
    a = O.fill_uniform_f32(M * K, 71, -1.0, 1.0).reshape(M, K); b = O.fill_uniform_f32(K * N, 72, -1.0, 1.0).reshape(K, N)
    nan = np.full((M, N), np.nan, np.float32)
    for sa, sb in ((1e-20, 1e-10), (1e+15, 3e+12), (1e-30, 1e+25), (7.0, 0.0), (1.0, 1.0)):
        aa, bb = (a * np.float32(sa)).astype(np.float32), (b * np.float32(sb)).astype(np.float32)
        got, *_ = run(L.PATH_F16X3, M, N, K, 1.0, aa, "row", bb, "row", 0.0, nan, "row")
        ex = aa.astype(np.float64) @ bb.astype(np.float64)
        bnd = (np.abs(aa).astype(np.float64) @ np.abs(bb).astype(np.float64)) * (3 * 2.0 ** -22 + 2e-6) + 1e-37
        assert np.isfinite(got).all() and (np.abs(got - ex) <= bnd).all(), (sa, sb)
    # one scale per row of A and per column of B: rows / columns of wildly different magnitude inside one matrix keep
    # their full per-product accuracy (block-scaled matrices), in every operand class
    rs_ = (2.0 ** np.random.default_rng(5).integers(-40, 40, M)).astype(np.float32)
    cs_ = (2.0 ** np.random.default_rng(6).integers(-40, 40, N)).astype(np.float32)
    aw, bw = (a * rs_[:, None]).astype(np.float32), (b * cs_[None, :]).astype(np.float32)
    ex = aw.astype(np.float64) @ bw.astype(np.float64)
    per = np.abs(aw).astype(np.float64) @ np.abs(bw).astype(np.float64) * (3 * 2.0 ** -22 + 2e-6)
    for la, lb in (("row", "row"), ("col", "col"), ("padded", "both2"), ("negrow", "misaligned")):
        got, *_ = run(L.PATH_F16X3, M, N, K, 1.0, aw, la, bw, lb, 0.0, nan, "row")
        assert np.isfinite(got).all() and (np.abs(got - ex) <= per).all(), (la, lb)
    # entries far below the maximum of their OWN row keep absolute precision (2^-39 of that maximum): the bound becomes
    # relative to max_k |a_ik| * sum_k |b_kj|, the row-norm error model of a blocked GEMM (documented domain of the mode)
    ai = a.copy(); ai[:, ::2] *= np.float32(2.0 ** -30)
    got, *_ = run(L.PATH_F16X3, M, N, K, 1.0, ai, "row", b, "row", 0.0, nan, "row")
    ex = ai.astype(np.float64) @ b.astype(np.float64)
    per = np.abs(ai).astype(np.float64) @ np.abs(b).astype(np.float64) * (3 * 2.0 ** -22 + 2e-6)
    assert (np.abs(got - ex) <= per + np.abs(ai).max(1)[:, None] * np.abs(b).astype(np.float64).sum(0)[None, :] * 2.0 ** -36).all()


@pytest.mark.skipif(EMU, reason="too large for the CPU build")
@pytest.mark.parametrize("path", MODES, ids=mode_ids)
def test_large_square_agrees_with_the_tf32x3_mode(path):
    """4096^3 (BASELINE.json configs[1] shape) on U(0,1): against the independent fp32-faithful mode TF32X3 (other operand
    pieces, other MMA kind); and the transposed-A layout of configs[2]"""
    import torch
    n = 4096
    a = torch.empty(n * n, device="cuda"); b = torch.empty(n * n, device="cuda")
    L.fill_uniform_f32(a, n * n, 61, 0.0, 1.0); L.fill_uniform_f32(b, n * n, 62, 0.0, 1.0)
    c1 = torch.full((n, n), float("nan"), device="cuda"); c2 = torch.full((n, n), float("nan"), device="cuda")
    for (rsa, csa) in ((n, 1), (1, n)):
        L.gemm_strided(n, n, n, 1.0, a, rsa, csa, b, n, 1, 0.0, c1, n, 1, path=path)
        L.gemm_strided(n, n, n, 1.0, a, rsa, csa, b, n, 1, 0.0, c2, n, 1, path=L.PATH_TF32X3)
        torch.cuda.synchronize()
        rel = ((c1 - c2).abs() / c2.abs()).max().item()
        assert rel < 1e-5, rel
