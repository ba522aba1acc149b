"""Scenarios run by tests/test_emulated_library.py in a subprocess whose LASER_B200_LIB points at the
host-emulated build of the WHOLE library (tests/emu_build.py: build_capi_host_emu).  "Device" memory is
host memory (numpy arrays wrapped in DevPtr); everything else is the ordinary Python mirror.  Each
scenario asserts and prints `OK <name> <count>`."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import laser_b200 as L  # noqa: E402
import oracle as O  # noqa: E402
from util import LAYOUTS, embed, f32_to_bf16_bits, bf16_bits_to_f32  # noqa: E402

NAME = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64",
        np.dtype(np.uint16): "bf16"}


def D(arr, off=0):
    """numpy array (kept alive by the caller) -> device pointer at element `off`"""
    return L.DevPtr(arr.ctypes.data + off * arr.itemsize, NAME[arr.dtype])


def rnd(shape, seed, lo=0.0, hi=1.0):
    return O.fill_uniform_f32(int(np.prod(shape)), seed, lo, hi).reshape(shape)


def ref_gemm(M, N, K, alpha, a, b, beta, c0):
    ref = c0.copy()
    O.gemm_strided(M, N, K, alpha, a, K, 1, b, N, 1, beta, ref, N, 1)
    return ref


TOL = {L.PATH_AUTO: 1e-5, L.PATH_F16X3: 1e-5, L.PATH_TF32X3: 1e-5, L.PATH_TF32X1: 3e-3, L.PATH_SIMT: 0.0}


def dispatch_and_modes():
    n = 0
    for (M, N, K) in ((64, 64, 64), (200, 300, 150), (130, 40, 70), (300, 9, 33)):
        a, b, c0 = rnd((M, K), 1), rnd((K, N), 2), rnd((M, N), 3)
        for path in (L.PATH_AUTO, L.PATH_TF32X3, L.PATH_TF32X1, L.PATH_SIMT):
            for alpha, beta in ((1.0, 0.0), (0.5, -1.25)):
                c = c0.copy() if beta else np.full((M, N), np.nan, np.float32)
                L.gemm_strided(M, N, K, alpha, D(a), K, 1, D(b), N, 1, beta, D(c), N, 1, path=path)
                ref = ref_gemm(M, N, K, alpha, a, b, beta, c0)
                small = M * N * K <= 128 ** 3
                want_path = L.PATH_SIMT if (path == L.PATH_SIMT or (path == L.PATH_AUTO and small)) else \
                    (L.PATH_F16X3 if path == L.PATH_AUTO else path)
                assert L.last_path() == want_path, (L.last_path(), want_path, M, N, K, path)
                tol = 0.0 if want_path == L.PATH_SIMT and alpha == 1.0 else max(TOL[path], 3e-7)
                err = np.abs(c - ref).max() / np.abs(ref).max()
                assert err <= tol, (err, tol, M, N, K, path, alpha, beta)
                n += 1
    print("OK dispatch_and_modes", n)


def strided_operands():
    """every operand class: K-major / MN-major TMA, general (gathered) -- and C of any strides"""
    M, N, K = 150, 140, 100
    a, b, c0 = rnd((M, K), 4), rnd((K, N), 5), rnd((M, N), 6)
    n = 0
    for la, lb, lc in [(x, "row", "row") for x in LAYOUTS] + [("row", x, "row") for x in LAYOUTS] + [("col", "col", x) for x in LAYOUTS]:
        A, oa, rsa, csa = embed(a, la); B, ob, rsb, csb = embed(b, lb); C, oc, rsc, csc = embed(c0, lc)
        Cref = C.copy()
        O.gemm_strided(M, N, K, 1.0, A[oa:], rsa, csa, B[ob:], rsb, csb, 2.0, Cref[oc:], rsc, csc)
        L.gemm_strided(M, N, K, 1.0, D(A, oa), rsa, csa, D(B, ob), rsb, csb, 2.0, D(C, oc), rsc, csc)   # AUTO -> tensor cores
        assert L.last_path() == L.PATH_F16X3
        idx = oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc
        assert np.abs(C[idx] - Cref[idx]).max() <= 1e-5 * np.abs(Cref[idx]).max(), (la, lb, lc)
        mask = np.ones(C.size, bool); mask[idx.reshape(-1)] = False
        assert np.array_equal(C[mask], Cref[mask]), (la, lb, lc)       # nothing outside the view is touched
        n += 1
    print("OK strided_operands", n)


def host_entry():
    """host-pointer (drop-in) entry: the staged path and, for M >= 2048, the pipelined row-panel path
    (panel geometry from LASER_B200_PANEL_ROWS / LASER_B200_PANEL_TAPER)"""
    n = 0
    for (M, N, K) in ((300, 70, 90), (2048, 24, 64), (2600, 16, 40)):
        a, b = rnd((M, K), 7), rnd((K, N), 8)
        c = np.full((M, N), np.nan, np.float32)
        L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1)
        ref = ref_gemm(M, N, K, 1.0, a, b, 0.0, np.zeros((M, N), np.float32))
        assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max(), (M, N, K)
        n += 1
    # beta != 0 on dense C: still the pipelined path, the old panel of C is uploaded next to its panel of A
    M, N, K = 2304, 20, 48
    a, b, c0 = rnd((M, K), 30), rnd((K, N), 31), rnd((M, N), 32)
    c = c0.copy()
    n0 = L.launch_count()
    L.gemm_strided(M, N, K, 0.5, a, K, 1, b, N, 1, -1.25, c, N, 1)
    assert L.launch_count() - n0 >= 1 + 2 * 2, L.launch_count() - n0     # prepare B once, then (prepare A, GEMM) per row panel
    ref = ref_gemm(M, N, K, 0.5, a, b, -1.25, c0)
    assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max()
    n += 1
    # padded C rows (span not dense) take the staged path even for large M
    M, N, K = 2100, 12, 40
    a, b = rnd((M, K), 9), rnd((K, N), 10)
    cbuf = np.full((M, N + 3), -7.0, np.float32); c0 = rnd((M, N), 11); cbuf[:, :N] = c0
    L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 1.0, cbuf, N + 3, 1)
    ref = ref_gemm(M, N, K, 1.0, a, b, 1.0, c0)
    assert np.abs(cbuf[:, :N] - ref).max() <= 1e-5 * np.abs(ref).max() and np.all(cbuf[:, N:] == -7.0)
    print("OK host_entry", n + 1)


def prepacked():
    n = 0
    for (M, N, K) in ((300, 520, 200), (100, 36, 77)):
        a, b, c0 = rnd((M, K), 12), rnd((K, N), 13), rnd((M, N), 14)
        pa = L.alloc_packed(L.gemm_prepackA_mem_required(M, N, K)); pb = L.alloc_packed(L.gemm_prepackB_mem_required(M, N, K))
        at = np.ascontiguousarray(a.T)                                  # A given transposed: rowStride 1, colStride M
        L.gemm_prepackA(pa, M, N, K, D(at), 1, M)
        L.gemm_prepackB(pb, M, N, K, D(b), N, 1)
        ref = ref_gemm(M, N, K, 0.5, a, b, 2.0, c0)
        c = c0.copy()
        L.gemm_packed(M, N, K, 0.5, pa, pb, 2.0, D(c), N, 1)
        assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max()
        c2 = c0.copy()
        L.gemm_packedB(M, N, K, 0.5, D(a), K, 1, pb, 2.0, D(c2), N, 1)
        assert np.array_equal(c2, c)                                      # same prepared operands, same kernel
        c3 = c0.copy()
        L.gemm_strided(M, N, K, 0.5, D(a), K, 1, D(b), N, 1, 2.0, D(c3), N, 1, path=L.PATH_F16X3)
        assert np.array_equal(c3, c)                                      # and the unpacked call agrees bit for bit
        n += 1
    print("OK prepacked", n)


def split_k():
    """few output tiles, long K (run with a many-SM device model): (tile, K-split) units + the reduce kernel"""
    M, N, K = 120, 200, 1100
    a, b, c0 = rnd((M, K), 15, -1, 1), rnd((K, N), 16, -1, 1), rnd((M, N), 17)
    n0 = L.launch_count()
    c = c0.copy()
    L.gemm_strided(M, N, K, 0.5, D(a), K, 1, D(b), N, 1, -1.0, D(c), N, 1, path=L.PATH_F16X3)
    launches = L.launch_count() - n0
    ref = ref_gemm(M, N, K, 0.5, a, b, -1.0, c0)
    assert np.abs(c - ref).max() <= 2e-5 * np.abs(ref).max()
    assert launches == 5, launches                                        # prepare A (fused), abs-max B, split B, GEMM, reduce
    os.environ  # (LASER_B200_SPLITK=0 variant is a separate process)
    print("OK split_k", launches)


def no_split_k():
    M, N, K = 120, 200, 1100
    a, b = rnd((M, K), 15, -1, 1), rnd((K, N), 16, -1, 1)
    n0 = L.launch_count()
    c = np.zeros((M, N), np.float32)
    L.gemm_strided(M, N, K, 1.0, D(a), K, 1, D(b), N, 1, 0.0, D(c), N, 1, path=L.PATH_F16X3)
    assert L.launch_count() - n0 == 4
    ref = ref_gemm(M, N, K, 1.0, a, b, 0.0, c * 0)
    assert np.abs(c - ref).max() <= 2e-5 * np.abs(ref).max()
    print("OK no_split_k 4")


def other_types():
    n = 0
    M, N, K = 70, 50, 300
    rng = np.random.default_rng(0)
    for dt, big in ((np.float64, None), (np.int32, 2 ** 31 - 5), (np.int64, 2 ** 62)):
        if big is None:
            a, b, c0 = rng.random((M, K)), rng.random((K, N)), rng.random((M, N))
            al, be = 1.0, 1.0
        else:
            a = rng.integers(-big, big, (M, K), dtype=dt); b = rng.integers(-big, big, (K, N), dtype=dt); c0 = rng.integers(-9, 9, (M, N), dtype=dt)
            al, be = 3, -2
        ref = c0.copy(); O.gemm_strided(M, N, K, al, a, K, 1, b, N, 1, be, ref, N, 1)
        c = c0.copy(); L.gemm_strided(M, N, K, al, D(a), K, 1, D(b), N, 1, be, D(c), N, 1)
        assert np.array_equal(c, ref), dt
        ch = c0.copy(); L.gemm_strided(M, N, K, al, a, K, 1, b, N, 1, be, ch, N, 1)       # host entry
        assert np.array_equal(ch, ref), dt
        n += 2
    # bf16 (tensor cores, fp32 accumulate, bf16 out)
    M, N, K = 200, 264, 100
    a, b, c0 = rnd((M, K), 18, -1, 1), rnd((K, N), 19, -1, 1), rnd((M, N), 20)
    ab, bb, cb = f32_to_bf16_bits(a).reshape(M, K), f32_to_bf16_bits(b).reshape(K, N), f32_to_bf16_bits(c0).reshape(M, N)
    ref = cb.copy(); O.gemm_strided(M, N, K, 1.0, ab, K, 1, bb, N, 1, 0.5, ref, N, 1, bf16=True)
    c = cb.copy(); L.gemm_strided(M, N, K, 1.0, D(ab), K, 1, D(bb), N, 1, 0.5, D(c), N, 1)
    assert L.last_path() == L.PATH_BF16
    assert np.abs(bf16_bits_to_f32(c) - bf16_bits_to_f32(ref)).max() <= 2.0 ** -7 * np.abs(bf16_bits_to_f32(ref)).max()
    print("OK other_types", n + 1)


def fused_and_skinny():
    M, N, K = 140, 270, 90
    a, b = rnd((M, K), 21, -1, 1), rnd((K, N), 22, -1, 1)
    bias = rnd((N,), 23, -1, 1)
    for path in (L.PATH_AUTO, L.PATH_SIMT):
        c = np.zeros((M, N), np.float32)
        L.gemm_strided_fused(M, N, K, 1.0, D(a), K, 1, D(b), N, 1, 0.0, D(c), N, 1, bias=D(bias), activation="relu", path=path)
        ref = np.maximum(ref_gemm(M, N, K, 1.0, a, b, 0.0, c * 0) + bias[None, :], 0)
        assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max(), path
    # skinny: N <= 4 and M >= 1024 -> warp-shuffle GEMV (three variants by alignment / size)
    n = 0
    for (M, K, NV, lda) in ((1024, 256, 1, 256), (1100, 300, 3, 301), (1024, 8192, 4, 8192)):
        A = rnd((M, lda), 24, -1, 1); b = rnd((K, NV), 25, -1, 1); c = np.zeros((M, NV), np.float32)
        L.gemm_strided(M, NV, K, 1.0, D(A), lda, 1, D(b), NV, 1, 0.0, D(c), NV, 1)
        ref = np.zeros((M, NV), np.float32); O.gemm_strided(M, NV, K, 1.0, A, lda, 1, b, NV, 1, 0.0, ref, NV, 1)
        assert np.abs(c - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), (M, K, NV)
        n += 1
    print("OK fused_and_skinny", n + 2)


def tensors():
    a = rnd((90, 40), 26); b = rnd((40, 60), 27)
    A, B = L.toTensor(a), L.toTensor(b)
    C = L.matmul(A, B)
    ref = ref_gemm(90, 60, 40, 1.0, a, b, 0.0, np.zeros((90, 60), np.float32))
    assert np.array_equal(C.to_numpy(), ref)                              # 90*60*40 <= 128^3: exact kernel
    Ct = L.newTensor([60, 90])
    L.matmul(B.transpose(), A.transpose(), Ct)                            # (AB)^T = B^T A^T on transposed views
    assert np.array_equal(Ct.to_numpy(), ref.T)
    buf = np.zeros(1000, np.float32)
    L.fill_uniform_f32(D(buf), 1000, 42, -0.1, 0.1)
    assert np.array_equal(buf, O.fill_uniform_f32(1000, 42, -0.1, 0.1))
    print("OK tensors 3")


def _two_piece_mode(PATH, name, per_product, gate, layouts_tol):
    """the default fp32 mode: two 16-bit pieces per operand, three passes of the 16-bit tensor-core kernel with fp32 output.
    Worst case 3 * 2^-22 of sum |a||b| from the dropped l*l' and remainder terms; the errors are random-signed, so on these
    (fixed, seeded) inputs a four times tighter bar holds with margin and catches a wrong pass order or a lost piece."""
    n = 0
    # (a) contiguous, sizes around the tile / accumulation-block boundaries, both scalings, both distributions
    for (M, N, K) in ((200, 300, 150), (130, 40, 70), (257, 260, 129), (300, 9, 333)):
        for lo, hi in ((0.0, 1.0), (-0.1, 0.1)):
            a, b, c0 = rnd((M, K), 41, lo, hi), rnd((K, N), 42, lo, hi), rnd((M, N), 43, lo, hi)
            for alpha, beta in ((1.0, 0.0), (0.5, -1.25)):
                c = c0.copy() if beta else np.full((M, N), np.nan, np.float32)
                L.gemm_strided(M, N, K, alpha, D(a), K, 1, D(b), N, 1, beta, D(c), N, 1, path=PATH)
                assert L.last_path() == PATH
                exact = alpha * (a.astype(np.float64) @ b.astype(np.float64)) + beta * c0.astype(np.float64)
                bound = abs(alpha) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)) * (per_product + 2e-6) \
                    + np.abs(exact) * 2e-6 + 1e-30
                assert (np.abs(c - exact) <= bound).all(), (float((np.abs(c - exact) / bound).max()), M, N, K, alpha, beta)
                if lo == 0.0:   # positive data: the north-star gate, max elementwise relative error vs the reference
                    ref = ref_gemm(M, N, K, alpha, a, b, beta, c0)
                    if beta == 0.0:
                        assert (np.abs(c - ref) / np.abs(ref)).max() < gate, (M, N, K)
                n += 1
    # (b) every operand class (K-major / MN-major split kernel, gathered general strides) and C of any strides
    M, N, K = 150, 140, 100
    a, b, c0 = rnd((M, K), 44), rnd((K, N), 45), rnd((M, N), 46)
    L.set_f32_mode(PATH)
    assert L.get_f32_mode() == PATH
    for la, lb, lc in [(x, "row", "row") for x in LAYOUTS] + [("row", x, "row") for x in LAYOUTS] + [("col", "col", x) for x in LAYOUTS]:
        A, oa, rsa, csa = embed(a, la); B, ob, rsb, csb = embed(b, lb); C, oc, rsc, csc = embed(c0, lc)
        Cref = C.copy()
        O.gemm_strided(M, N, K, 1.0, A[oa:], rsa, csa, B[ob:], rsb, csb, 2.0, Cref[oc:], rsc, csc)
        L.gemm_strided(M, N, K, 1.0, D(A, oa), rsa, csa, D(B, ob), rsb, csb, 2.0, D(C, oc), rsc, csc)   # AUTO -> the selected mode
        assert L.last_path() == PATH
        idx = oc + np.arange(M)[:, None] * rsc + np.arange(N)[None, :] * csc
        assert np.abs(C[idx] - Cref[idx]).max() <= layouts_tol * np.abs(Cref[idx]).max(), (la, lb, lc)
        mask = np.ones(C.size, bool); mask[idx.reshape(-1)] = False
        assert np.array_equal(C[mask], Cref[mask]), (la, lb, lc)
        n += 1
    # (c) host-pointer entry in this mode (staged path), and back to the default
    M, N, K = 2100, 24, 72
    a, b = rnd((M, K), 47), rnd((K, N), 48)
    c = np.full((M, N), np.nan, np.float32)
    L.gemm_strided(M, N, K, 1.0, a, K, 1, b, N, 1, 0.0, c, N, 1)
    ref = ref_gemm(M, N, K, 1.0, a, b, 0.0, c)
    assert (np.abs(c - ref) / np.abs(ref)).max() < gate
    L.set_f32_mode(L.PATH_F16X3)
    n += 1
    return n


def f16x3():
    """two fp16 pieces of the power-of-two-scaled operands (device-side abs-max), fp16 flavour of the kernel whose
    epilogue undoes the scales: accuracy of tf32x3, and fp32's range although fp16 has 5 exponent bits"""
    n = _two_piece_mode(L.PATH_F16X3, "f16x3", 3 * 2.0 ** -22, 3e-6, 3e-6)
    M, N, K = 200, 130, 96
    a, b = rnd((M, K), 61, -1.0, 1.0), rnd((K, N), 62, -1.0, 1.0)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    absab = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    # operands far outside fp16's range, in both directions, and a zero operand (abs-max word 0: no scaling)
    for sa, sb in ((1e-20, 1e-10), (1e+15, 3e+12), (1e-30, 1e+25), (7.0, 0.0)):
        aa, bb = (a * np.float32(sa)).astype(np.float32), (b * np.float32(sb)).astype(np.float32)
        c = np.full((M, N), np.nan, np.float32)
        L.gemm_strided(M, N, K, 1.0, D(aa), K, 1, D(bb), N, 1, 0.0, D(c), N, 1, path=L.PATH_F16X3)
        ex = aa.astype(np.float64) @ bb.astype(np.float64)
        bound = (np.abs(aa).astype(np.float64) @ np.abs(bb).astype(np.float64)) * (3 * 2.0 ** -22 + 2e-6) + 1e-37
        assert np.isfinite(c).all() and (np.abs(c - ex) <= bound).all(), (sa, sb, float((np.abs(c - ex) / bound).max()))
        n += 1
    # one scale per row of A and per column of B: rows / columns of wildly different magnitude inside one matrix keep
    # their full per-product accuracy (block-scaled matrices) ...
    rs_ = (2.0 ** np.random.default_rng(5).integers(-40, 40, M)).astype(np.float32)
    cs_ = (2.0 ** np.random.default_rng(6).integers(-40, 40, N)).astype(np.float32)
    aw, bw = (a * rs_[:, None]).astype(np.float32), (b * cs_[None, :]).astype(np.float32)
    for (la, lb) in (("row", "row"), ("col", "col"), ("padded", "both2"), ("negrow", "misaligned")):     # K-major, MN-major, padded, gathered
        A, oa, rsa, csa = embed(aw, la); B, ob, rsb, csb = embed(bw, lb)
        c = np.full((M, N), np.nan, np.float32)
        L.gemm_strided(M, N, K, 1.0, D(A, oa), rsa, csa, D(B, ob), rsb, csb, 0.0, D(c), N, 1, path=L.PATH_F16X3)
        ex = aw.astype(np.float64) @ bw.astype(np.float64)
        per = np.abs(aw).astype(np.float64) @ np.abs(bw).astype(np.float64) * (3 * 2.0 ** -22 + 2e-6)
        assert np.isfinite(c).all() and (np.abs(c - ex) <= per).all(), (la, lb, float((np.abs(c - ex) / per).max()))
        n += 1
    # ... while entries far below the maximum of their OWN row keep absolute precision (2^-39 of that maximum): the
    # error bound becomes relative to max_k |a_ik| * sum_k |b_kj|, the row-norm model of a blocked GEMM
    ai = a.copy(); ai[:, ::2] *= np.float32(2.0 ** -30)
    c = np.full((M, N), np.nan, np.float32)
    L.gemm_strided(M, N, K, 1.0, D(ai), K, 1, D(b), N, 1, 0.0, D(c), N, 1, path=L.PATH_F16X3)
    ex = ai.astype(np.float64) @ b.astype(np.float64)
    assert (np.abs(c - ex) <= np.abs(ai).astype(np.float64) @ np.abs(b).astype(np.float64) * (3 * 2.0 ** -22 + 2e-6)
            + np.abs(ai).max(1)[:, None] * np.abs(b).astype(np.float64).sum(0)[None, :] * 2.0 ** -36).all()
    # two calls in a row with different ranges: the abs-max words are per call
    c2 = np.full((M, N), np.nan, np.float32)
    L.gemm_strided(M, N, K, 1.0, D(a), K, 1, D(b), N, 1, 0.0, D(c2), N, 1, path=L.PATH_F16X3)
    assert (np.abs(c2 - exact) <= absab * (3 * 2.0 ** -22 + 2e-6)).all()
    print("OK f16x3", n + 2)


def rowsharded():
    """the multi-GPU entry points behind the C ABI (capi_multi.inc) on LASER_B200_EMU_DEVICES emulated devices with a stand-in
    NCCL (tests/emu/fake_nccl.c): the host-pointer entry (uploads, one broadcast of B inside an NCCL group, every device its
    row panel, downloads) and the per-rank device entry driven the way a single-process caller must drive it"""
    import ctypes
    from laser_b200 import rowshard as RS
    ndev = int(os.environ["LASER_B200_EMU_DEVICES"])
    n = 0
    for (M, N, K, alpha, beta) in ((700, 140, 100, 1.0, 0.0), (1000, 64, 72, 0.5, -1.25), (200, 40, 64, 1.0, 2.0)):
        a, b, c0 = rnd((M, K), 70), rnd((K, N), 71), rnd((M, N), 72)
        c = c0.copy() if beta else np.full((M, N), np.nan, np.float32)
        RS.gemm_rowsharded_host(ndev, M, N, K, alpha, a, K, 1, b, N, 1, beta, c, N, 1)
        ref = ref_gemm(M, N, K, alpha, a, b, beta, c0)
        assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max(), (M, N, K)
        assert RS.partition_rows(M, ndev) == [RS.partition_rows_c(M, ndev, r) for r in range(ndev)]
        n += 1
    # per-rank device entry (one process or thread per GPU in real use; played here rank after rank, the root first, which
    # is the order the stream dependencies impose anyway): B valid on the root only, every rank its rows
    fake = ctypes.CDLL(os.environ["LASER_B200_NCCL_LIB"])
    cudart_set = L.lib().cudaSetDevice
    comms = RS.comm_init_all(ndev)
    assert [cm.rank for cm in comms] == list(range(ndev)) and all(cm.size == ndev for cm in comms)
    M, N, K = 900, 72, 80
    a, bfull, c0 = rnd((M, K), 73), rnd((K, N), 74), rnd((M, N), 75)
    parts = RS.partition_rows(M, ndev)
    Bs = [bfull.copy() if r == 1 else np.full((K, N), np.nan, np.float32) for r in range(ndev)]   # root = rank 1
    Cs = [c0[lo:hi].copy() for lo, hi in parts]
    calls0 = fake.fake_nccl_broadcast_calls()
    for r in [1] + [x for x in range(ndev) if x != 1]:
        lo, hi = parts[r]
        assert cudart_set(r) == 0
        RS.gemm_rowsharded_dev(comms[r], hi - lo, N, K, 0.5, D(a[lo:hi]) if hi > lo else None, K, 1, D(Bs[r]), N, 1, 1, -1.25,
                               D(Cs[r]) if hi > lo else None, N, 1, stream=1)
    cudart_set(0)
    assert fake.fake_nccl_broadcast_calls() - calls0 == ndev
    ref = ref_gemm(M, N, K, 0.5, a, bfull, -1.25, c0)
    for r, (lo, hi) in enumerate(parts):
        assert np.array_equal(Bs[r], bfull), r
        assert hi == lo or np.abs(Cs[r] - ref[lo:hi]).max() <= 1e-5 * np.abs(ref).max(), r     # (a rank may own no rows)
    # default fp32 mode, B row- or column-major, N >= 512: B travels PREPARED in column panels (capi_multi.inc:
    # rowshard_prepared) -- the root prepares and sends panel after panel, every rank multiplies its rows by each panel as it
    # arrives; B of the other ranks is not touched
    P = int(os.environ.get("LASER_B200_ROWSHARD_PANELS", "2"))
    if P >= 1:
        for row_major in (True, False):
            M, N, K = 300, 2304, 96
            a, bfull, c0 = rnd((M, K), 76), rnd((K, N), 77), rnd((M, N), 78)
            parts = RS.partition_rows(M, ndev)
            stored = bfull if row_major else np.ascontiguousarray(bfull.T)        # row-major [K][N] or column-major (= [N][K])
            rsb, csb = (N, 1) if row_major else (1, K)
            Bs = [stored.copy() if r == 0 else np.full(stored.shape, np.nan, np.float32) for r in range(ndev)]
            Cs = [c0[lo:hi].copy() for lo, hi in parts]
            calls0 = fake.fake_nccl_broadcast_calls()
            for r in range(ndev):
                lo, hi = parts[r]
                assert cudart_set(r) == 0
                RS.gemm_rowsharded_dev(comms[r], hi - lo, N, K, 0.5, D(a[lo:hi]) if hi > lo else None, K, 1, D(Bs[r]), rsb, csb, 0,
                                       -1.25, D(Cs[r]) if hi > lo else None, N, 1, stream=1)
            cudart_set(0)
            W = -(-(-(-N // min(P, N // 1024)) ) // 256) * 256
            panels = -(-N // W)
            per_rank = (1 + panels) if row_major else 2 * panels
            assert fake.fake_nccl_broadcast_calls() - calls0 == ndev * per_rank, (fake.fake_nccl_broadcast_calls() - calls0, panels)
            ref = ref_gemm(M, N, K, 0.5, a, bfull, -1.25, c0)
            assert np.array_equal(Bs[0], stored)
            for r, (lo, hi) in enumerate(parts):
                assert r == 0 or np.isnan(Bs[r]).all(), r                       # B of the other ranks: not an output
                assert hi == lo or np.abs(Cs[r] - ref[lo:hi]).max() <= 2e-5 * np.abs(ref).max(), (row_major, r)
            n += 1
    for cm in comms:
        cm.destroy()
    print("OK rowsharded", n + 1)


def lifecycle():
    """init / shutdown / re-init: workspaces, staging buffers and the layer workspace are released and rebuilt"""
    a, b = rnd((300, 200), 40), rnd((200, 260), 41)
    ref = ref_gemm(300, 260, 200, 1.0, a, b, 0.0, np.zeros((300, 260), np.float32))
    for _ in range(2):
        L.init()
        c = np.zeros((300, 260), np.float32)
        L.gemm_strided(300, 260, 200, 1.0, a, 200, 1, b, 260, 1, 0.0, c, 260, 1)          # host entry: staging buffers
        assert np.abs(c - ref).max() <= 1e-5 * np.abs(ref).max()
        out = np.zeros((1, 1, 4, 4), np.float32)
        L.conv2d_im2col(out, np.ones((1, 1, 4, 4), np.float32), (1, 1, 4, 4), np.ones((1, 1, 3, 3), np.float32), (1, 1, 3, 3), (1, 1), (1, 1))
        assert out[0, 0, 1, 1] == 9.0 and out[0, 0, 0, 0] == 4.0                           # host conv: layer workspace
        L.shutdown()
    print("OK lifecycle 2")


if __name__ == "__main__":
    globals()[sys.argv[1]]()
