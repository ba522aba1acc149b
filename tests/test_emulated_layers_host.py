"""CPU-only: the HOST side of the layer entry points (laser_b200/csrc/capi_layers.inc -- argument
checks, convolution geometry, the im2col + batched-GEMM loop with workspace chunking, host-pointer
staging, view handling of copyFrom / forEach) compiled for the CPU (tests/emu/capi_layers_emu.cpp) and
called through the same ctypes signatures as the real library.  "Device" memory is host memory, kernels
run on host threads, every GEMM goes through the emulated exact kernel, so convolutions are compared
with the oracle bit for bit."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle as O
from emu_build import build_emu
from laser_b200._capi import SIGNATURES, TensorView

HERE = os.path.dirname(os.path.abspath(__file__))
LAYER_SYMBOLS = [n for n in SIGNATURES if any(k in n for k in ("transpose2D", "nchw", "nhwc", "conv2d", "im2col", "batched",
                                                               "copy_views", "foreach_views"))]
i64 = ctypes.c_int64
EINVAL, EUNSUPPORTED = 1, 5


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(build_emu("capi_layers_emu", ["capi_layers.inc", "layers.cuh", "gemm_simt.cuh", "../../include/laser_b200.h"]))
    for name in LAYER_SYMBOLS:
        fn = getattr(L, name)                      # every layer symbol of the header exists in this build too
        fn.restype, fn.argtypes = SIGNATURES[name]
    L.emu_last_error.restype = ctypes.c_char_p
    L.emu_launch_count.restype = i64
    return L


def p(a, off=0):
    return ctypes.c_void_p(a.ctypes.data + off * a.itemsize)


def i4(t):
    return (i64 * 4)(*t)


def i2(t):
    return (i64 * 2)(*t)


def view(arr, shape, strides, offset=0, dtype=0):
    v = TensorView()
    v.rank = len(shape); v.dtype = dtype
    for i, (n, s) in enumerate(zip(shape, strides)):
        v.shape[i] = n; v.strides[i] = s
    v.offset = offset
    v.storage = arr.ctypes.data
    return v


def conv_ref(inp, ishape, ker, kshape, padding, strides):
    """Oracle convolution.  For 1x1 kernels with a stride or padding the reference's im2col shortcut
    (conv2d_im2col.nim:121,145-149) reads the image in place and is wrong; the product deliberately
    goes through im2col there (include/laser_b200.h), so the expectation is im2col + the GEMM oracle."""
    if kshape[2] * kshape[3] != 1 or (tuple(strides) == (1, 1) and tuple(padding) == (0, 0)):
        return O.conv2d_im2col(inp, ishape, ker, kshape, padding, strides)
    o = O.conv2d_out_shape(ishape, kshape, padding, strides)
    M, K, N = kshape[0], ishape[1], o[2] * o[3]
    out = np.zeros((ishape[0], M, N), np.float32)
    kmat = np.ascontiguousarray(ker, np.float32).reshape(M, K)
    for n in range(ishape[0]):
        ws = np.ascontiguousarray(O.im2col(np.ascontiguousarray(inp[n]), ishape, kshape, padding, strides))
        O.gemm_strided(M, N, K, 1.0, kmat, K, 1, ws, N, 1, 0.0, out[n], N, 1)
    return out.reshape(o)


# ---- transposes ---------------------------------------------------------------------------------
@pytest.mark.parametrize("esz,dt", [(1, np.uint8), (2, np.uint16), (4, np.float32), (8, np.float64)])
def test_transposes_device_and_host_entries(lib, esz, dt):
    N, C, H, W = 3, 5, 6, 7
    x = (np.arange(N * C * H * W) % 251).astype(dt)
    y = np.zeros_like(x); z = np.zeros_like(x); hy = np.zeros_like(x)
    assert lib.laser_b200_nchw2nhwc_dev(p(y), p(x), N, C, H, W, esz, None) == 0
    assert np.array_equal(y.reshape(N, H, W, C), x.reshape(N, C, H, W).transpose(0, 2, 3, 1))
    assert lib.laser_b200_nhwc2nchw_dev(p(z), p(y), N, C, H, W, esz, None) == 0
    assert np.array_equal(z, x)
    assert lib.laser_b200_nchw2nhwc(p(hy), p(x), N, C, H, W, esz) == 0           # host entry: staged, synchronous
    assert np.array_equal(hy, y)
    t = np.zeros(C * H, dt)
    assert lib.laser_b200_transpose2D_copy_dev(p(t), p(x), C, H, esz, None) == 0
    assert np.array_equal(t.reshape(H, C), x[:C * H].reshape(C, H).T)
    t2 = np.zeros(2 * C * H, dt)
    assert lib.laser_b200_transpose2D_batched(p(t2), p(x), 2, C, H, esz) == 0
    assert np.array_equal(t2.reshape(2, H, C), x[:2 * C * H].reshape(2, C, H).transpose(0, 2, 1))


def test_transpose_argument_checks(lib):
    a = np.zeros(16, np.float32); b = np.ones(16, np.float32)
    assert lib.laser_b200_transpose2D_copy_dev(p(a), p(a), 4, 4, 4, None) == EINVAL        # aliasing
    assert b"alias" in lib.emu_last_error()
    assert lib.laser_b200_transpose2D_copy_dev(p(a), p(b), -1, 4, 4, None) == EINVAL
    assert lib.laser_b200_transpose2D_copy_dev(p(a), p(b), 4, 4, 3, None) == EUNSUPPORTED  # element size
    assert lib.laser_b200_transpose2D_copy_dev(None, p(b), 4, 4, 4, None) == EINVAL
    assert lib.laser_b200_transpose2D_copy_dev(p(a), p(b), 0, 4, 4, None) == 0 and np.all(a == 0)   # nothing to do


# ---- convolution ----------------------------------------------------------------------------------
def conv_cases():
    with open(os.path.join(HERE, "golden", "conv2d_known_answer.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", conv_cases(), ids=lambda c: c["src"])
def test_conv2d_known_answer_device_and_host(lib, case):
    inp = np.array(case["input"], np.float32); ker = np.array(case["kernel"], np.float32)
    tgt = np.array(case["target"], np.float32)
    ish, ksh, pad, st = case["ishape"], case["kshape"], case["padding"], case["strides"]
    osh = (i64 * 4)()
    assert lib.laser_b200_conv2d_out_shape(i4(ish), i4(ksh), i2(pad), i2(st), osh) == 0 and tuple(osh) == tgt.shape
    per = lib.laser_b200_im2col_workspace_size(i4(ish), i4(ksh), i2(pad), i2(st))
    assert per == O.im2col_workspace_size(ish, ksh, pad, st)
    out = np.full(tgt.shape, 99.0, np.float32); ws = np.zeros(per, np.float32)
    assert lib.laser_b200_conv2d_im2col_f32_dev(p(out), p(inp), i4(ish), p(ker), i4(ksh), i2(pad), i2(st), p(ws), 1, 0, None) == 0
    assert np.array_equal(out, tgt)
    hout = np.full(tgt.shape, 99.0, np.float32)
    assert lib.laser_b200_conv2d_im2col_f32(p(hout), p(inp), i4(ish), p(ker), i4(ksh), i2(pad), i2(st)) == 0
    assert np.array_equal(hout, tgt)


@pytest.mark.parametrize("ishape,kshape,padding,strides,ws_images", [
    ((2, 3, 9, 11), (4, 3, 3, 3), (0, 0), (1, 1), 1),
    ((5, 2, 8, 8), (5, 2, 3, 3), (1, 1), (2, 2), 2),            # batch not a multiple of the workspace
    ((3, 4, 7, 10), (3, 4, 1, 1), (0, 0), (1, 1), 1),           # 1x1, unit stride: no im2col, image read in place
    ((2, 4, 9, 9), (3, 4, 1, 1), (1, 1), (2, 2), 2),            # strided / padded 1x1 goes through im2col
    ((4, 2, 6, 6), (2, 2, 3, 3), (1, 1), (1, 1), 9),            # workspace larger than the batch
])
def test_conv2d_matches_oracle_bit_for_bit(lib, ishape, kshape, padding, strides, ws_images):
    inp = O.fill_uniform_f32(int(np.prod(ishape)), 31, -1, 1).reshape(ishape)
    ker = O.fill_uniform_f32(int(np.prod(kshape)), 32, -1, 1).reshape(kshape)
    ref = conv_ref(inp, ishape, ker, kshape, padding, strides)
    out = np.full(ref.shape, np.nan, np.float32)
    per = lib.laser_b200_im2col_workspace_size(i4(ishape), i4(kshape), i2(padding), i2(strides))
    ws = np.full(ws_images * per + 4, -3.0, np.float32)
    n0 = lib.emu_launch_count()
    assert lib.laser_b200_conv2d_im2col_f32_dev(p(out), p(inp), i4(ishape), p(ker), i4(kshape), i2(padding), i2(strides), p(ws),
                                                ws_images, 0, None) == 0
    assert np.array_equal(out, ref)
    assert np.all(ws[ws_images * per:] == -3.0)
    in_place = kshape[2] * kshape[3] == 1 and strides == (1, 1) and padding == (0, 0)
    chunks = 0 if in_place else -(-ishape[0] // ws_images)
    assert lib.emu_launch_count() - n0 in (chunks + chunks, chunks + ishape[0], ishape[0], 1)   # im2col launches + GEMM launches
    hout = np.zeros_like(out)
    assert lib.laser_b200_conv2d_im2col_f32(p(hout), p(inp), i4(ishape), p(ker), i4(kshape), i2(padding), i2(strides)) == 0
    assert np.array_equal(hout, ref)


def test_conv2d_argument_checks(lib):
    x = np.zeros(64, np.float32)
    osh = (i64 * 4)()
    assert lib.laser_b200_conv2d_out_shape(i4((1, 1, 4, 4)), i4((1, 1, 3, 3)), i2((0, 0)), i2((4, 1)), osh) == EINVAL   # stride >= extent
    assert lib.laser_b200_conv2d_out_shape(i4((1, 1, 4, 4)), i4((1, 1, 7, 3)), i2((0, 0)), i2((1, 1)), osh) == EINVAL   # kernel > image
    assert lib.laser_b200_im2col_workspace_size(i4((1, 1, 4, 4)), i4((1, 1, 3, 3)), i2((0, 0)), i2((4, 1))) == -1
    args = (i4((1, 1, 4, 4)), p(x), i4((1, 2, 3, 3)), i2((1, 1)), i2((1, 1)))
    assert lib.laser_b200_conv2d_im2col_f32_dev(p(x), p(x), args[0], args[1], args[2], args[3], args[4], p(x), 1, 0, None) == EINVAL  # c_in
    ok = (i4((1, 1, 4, 4)), p(x), i4((1, 1, 3, 3)), i2((1, 1)), i2((1, 1)))
    assert lib.laser_b200_conv2d_im2col_f32_dev(p(x), p(x), ok[0], ok[1], ok[2], ok[3], ok[4], None, 1, 0, None) == EINVAL          # no workspace
    assert lib.laser_b200_conv2d_im2col_f32_dev(None, p(x), ok[0], ok[1], ok[2], ok[3], ok[4], p(x), 1, 0, None) == EINVAL
    empty = (i4((0, 1, 4, 4)), p(x), i4((1, 1, 3, 3)), i2((1, 1)), i2((1, 1)))
    assert lib.laser_b200_conv2d_im2col_f32_dev(p(x), p(x), empty[0], empty[1], empty[2], empty[3], empty[4], p(x), 1, 0, None) == 0  # empty batch


def test_im2col_several_images_one_launch(lib):
    ish, ksh, pad, st = (3, 2, 8, 8), (5, 2, 3, 3), (1, 1), (2, 2)
    inp = O.fill_uniform_f32(int(np.prod(ish)), 5, 1, 2).reshape(ish)
    per = O.im2col_workspace_size(ish, ksh, pad, st)
    ws = np.zeros(3 * per, np.float32)
    n0 = lib.emu_launch_count()
    assert lib.laser_b200_im2col_f32_dev(p(ws), p(inp), 3, i4(ish), i4(ksh), i2(pad), i2(st), None) == 0
    assert lib.emu_launch_count() - n0 == 1
    for b in range(3):
        assert np.array_equal(ws[b * per:(b + 1) * per], O.im2col(inp[b], ish, ksh, pad, st).reshape(-1))


# ---- batched GEMM ---------------------------------------------------------------------------------
def test_batched_dispatch(lib):
    batch, M, N, K = 6, 20, 24, 30                # below the 128^3 threshold: one launch for the whole batch
    A = O.fill_uniform_f32(batch * M * K, 1, -1, 1); B = O.fill_uniform_f32(K * N, 2, -1, 1)
    C = np.full(batch * M * N, np.nan, np.float32); ref = np.zeros(batch * M * N, np.float32)
    O.gemm_strided_batched(batch, M, N, K, 1.0, A, K, 1, M * K, B, N, 1, 0, 0.0, ref, N, 1, M * N)
    n0 = lib.emu_launch_count()
    assert lib.laser_b200_gemm_strided_batched_f32_dev(batch, M, N, K, 1.0, p(A), K, 1, M * K, p(B), N, 1, 0, 0.0, p(C), N, 1, M * N, 0, None) == 0
    assert lib.emu_launch_count() - n0 == 1 and np.array_equal(C, ref)
    batch, M, N, K = 3, 130, 129, 128             # above it: one GEMM call per problem
    A = O.fill_uniform_f32(batch * M * K, 3, -1, 1); B = O.fill_uniform_f32(batch * K * N, 4, -1, 1)
    C = np.zeros(batch * M * N, np.float32); ref = C.copy()
    O.gemm_strided_batched(batch, M, N, K, 1.0, A, K, 1, M * K, B, N, 1, K * N, 0.0, ref, N, 1, M * N)
    n0 = lib.emu_launch_count()
    assert lib.laser_b200_gemm_strided_batched_f32_dev(batch, M, N, K, 1.0, p(A), K, 1, M * K, p(B), N, 1, K * N, 0.0, p(C), N, 1, M * N, 0, None) == 0
    assert lib.emu_launch_count() - n0 == batch and lib.emu_last_requested_path() == 0 and np.array_equal(C, ref)
    assert lib.laser_b200_gemm_strided_batched_f32_dev(-1, M, N, K, 1.0, p(A), K, 1, 0, p(B), N, 1, 0, 0.0, p(C), N, 1, 0, 0, None) == EINVAL
    Ad = np.random.default_rng(0).random(2 * 9 * 7); Bd = np.random.default_rng(1).random(7 * 5); Cd = np.zeros(2 * 9 * 5); refd = Cd.copy()
    for b in range(2):
        O.gemm_strided(9, 5, 7, 1.0, Ad[b * 63:], 7, 1, Bd, 5, 1, 0.0, refd[b * 45:(b + 1) * 45], 5, 1)
    assert lib.laser_b200_gemm_strided_batched_f64_dev(2, 9, 5, 7, 1.0, p(Ad), 7, 1, 63, p(Bd), 5, 1, 0, 0.0, p(Cd), 5, 1, 45, None) == 0
    assert np.array_equal(Cd, refd)


# ---- copyFrom / forEach on views -----------------------------------------------------------------
def test_copy_views(lib):
    src = np.arange(40 * 60, dtype=np.float32); dst = np.zeros(60 * 40, np.float32)
    vs = view(src, (60, 40), (1, 60)); vd = view(dst, (60, 40), (40, 1))         # materialise the transpose
    assert lib.laser_b200_copy_views(ctypes.byref(vd), ctypes.byref(vs), None) == 0
    assert np.array_equal(dst.reshape(60, 40), src.reshape(40, 60).T)
    n0 = lib.emu_launch_count()
    dst2 = np.zeros(2400, np.float32)
    assert lib.laser_b200_copy_views(ctypes.byref(view(dst2, (40, 60), (60, 1))), ctypes.byref(view(src, (40, 60), (60, 1))), None) == 0
    assert lib.emu_launch_count() == n0 and np.array_equal(dst2, src)            # contiguous pair: plain copy, no kernel
    d64 = np.zeros(50, np.float64); s64 = np.arange(100, dtype=np.float64)
    assert lib.laser_b200_copy_views(ctypes.byref(view(d64, (5, 5), (10, 2), 0, 1)), ctypes.byref(view(s64, (5, 5), (20, 4), 1, 1)), None) == 0
    exp = np.zeros(50); exp.reshape(5, 10)[:, ::2] = s64[1:].reshape(-1)[:99].copy().reshape(-1)[np.arange(5)[:, None] * 20 + np.arange(5)[None, :] * 4]
    assert np.array_equal(d64, exp)
    assert lib.laser_b200_copy_views(ctypes.byref(view(dst, (60, 41), (41, 1))), ctypes.byref(vs), None) == EINVAL   # shape mismatch
    assert lib.laser_b200_copy_views(ctypes.byref(view(dst, (60, 40), (40, 1), 0, 1)), ctypes.byref(vs), None) == EINVAL  # dtype mismatch


def test_foreach_views(lib):
    x = np.arange(12, dtype=np.float32); y = np.full(12, 2.0, np.float32); z = np.full(12, 3.0, np.float32)
    vx, vy, vz = view(x, (3, 4), (4, 1)), view(y, (3, 4), (4, 1)), view(z, (3, 4), (1, 3))
    assert lib.laser_b200_foreach_views(6, ctypes.byref(vx), ctypes.byref(vx), ctypes.byref(vy), ctypes.byref(vz), 0.0, None) == 0
    assert np.array_equal(x, np.arange(12, dtype=np.float32) + 6)                # x += y * z in place
    o = np.zeros(12, np.float32)
    assert lib.laser_b200_foreach_views(1, ctypes.byref(view(o, (3, 4), (4, 1))), None, None, None, 2.5, None) == 0 and np.all(o == 2.5)
    assert lib.laser_b200_foreach_views(3, ctypes.byref(vx), ctypes.byref(vy), None, None, 0.0, None) == EINVAL     # missing operand
    assert lib.laser_b200_foreach_views(42, ctypes.byref(vx), None, None, None, 0.0, None) == EINVAL
    assert lib.laser_b200_foreach_views(3, ctypes.byref(vx), ctypes.byref(vy), ctypes.byref(view(z, (4, 3), (3, 1))), None, 0.0, None) == EINVAL
    oi = np.zeros(12, np.int32)
    assert lib.laser_b200_foreach_views(1, ctypes.byref(view(oi, (3, 4), (4, 1), 0, 2)), None, None, None, 1.0, None) == EUNSUPPORTED
