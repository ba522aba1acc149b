"""CPU-only: the kernels of laser_b200/csrc/layers.cuh (transposition, im2col, strided copy) run on
host threads (tests/emu/cuda_emu.h: one thread per CUDA thread, a barrier for __syncthreads) and are
compared with the oracle bit for bit.  This executes the product's kernel source -- index
arithmetic, bounds, shared-memory choreography, launch planning -- without a GPU; the GPU tests
(test_gpu_layers.py) then only have to confirm the same on the device."""
import ctypes

import numpy as np
import pytest

import oracle as O

from emu_build import build_emu

i64 = ctypes.c_int64


@pytest.fixture(scope="module")
def emu():
    so = build_emu("layers_emu", ["layers.cuh"])
    L = ctypes.CDLL(so)
    i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    L.emu_transpose_batched.restype = ci
    L.emu_transpose_batched.argtypes = [ci, vp, vp, i64, i64, i64, ci, ci]
    L.emu_im2col.restype = ci
    L.emu_im2col.argtypes = [vp, vp, i64, i64 * 11, ci]
    L.emu_copy_strided.restype = ci
    L.emu_copy_strided.argtypes = [ci, vp, vp, ci, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64), ci]
    P = ctypes.POINTER(i64)
    L.emu_foreach.restype = ci
    L.emu_foreach.argtypes = [ci, ci, vp, vp, vp, vp, ci, P, P, P, P, P, ctypes.c_double, ci]
    return L


def ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.float32, np.float64])
@pytest.mark.parametrize("N,NR,NC,grid", [
    (1, 64, 64, 0), (1, 128, 192, 3), (2, 68, 132, 2),      # multiples of 4: vector path, ragged tiles
    (1, 1, 1, 0), (1, 65, 63, 0), (3, 33, 70, 4), (1, 5, 257, 2),   # scalar path
    (1, 4, 1000, 0), (1, 1000, 4, 5),
])
def test_transpose_kernel(emu, dt, N, NR, NC, grid):
    src = (np.arange(N * NR * NC, dtype=np.int64) * 2654435761 % 65521).astype(dt)
    dst = np.full(N * NR * NC, 77, dt)
    v = emu.emu_transpose_batched(np.dtype(dt).itemsize, ptr(dst), ptr(src), N, NR, NC, grid, 0)
    assert v == (4 if NR % 4 == 0 and NC % 4 == 0 else 1)
    assert np.array_equal(dst.reshape(N, NC, NR), O.transpose2D_batched(src, N, NR, NC))
    if v == 4:   # the scalar kernel must agree on the same shape
        dst2 = np.zeros_like(dst)
        assert emu.emu_transpose_batched(np.dtype(dt).itemsize, ptr(dst2), ptr(src), N, NR, NC, grid, 1) == 1
        assert np.array_equal(dst2, dst)


def test_transpose_misaligned_base_takes_scalar_path(emu):
    N, NR, NC = 1, 64, 64
    buf = np.arange(NR * NC + 1, dtype=np.float32)
    src = buf[1:]                                   # 4 bytes off a 16-byte boundary
    dst = np.zeros(NR * NC, np.float32)
    assert emu.emu_transpose_batched(4, ptr(dst), ptr(src), N, NR, NC, 0, 0) == 1
    assert np.array_equal(dst.reshape(NC, NR), src.reshape(NR, NC).T)


IM2COL_CASES = [
    # ishape (images, C, H, W), kshape (c_out, c_in, kH, kW), padding, strides
    ((1, 1, 4, 4), (1, 1, 3, 3), (1, 1), (1, 1)),          # the reference's first known-answer geometry
    ((1, 3, 5, 5), (2, 3, 3, 3), (1, 1), (2, 2)),          # the second
    ((2, 3, 9, 11), (4, 3, 3, 3), (0, 0), (1, 1)),
    ((3, 2, 8, 8), (5, 2, 3, 3), (1, 1), (2, 2)),          # outHW = 16: float4 stores
    ((2, 1, 12, 6), (2, 1, 5, 2), (2, 1), (3, 3)),
    ((1, 2, 40, 36), (1, 2, 3, 3), (1, 1), (1, 1)),        # outHW = 1440 > 1024: several column chunks
    ((2, 5, 7, 7), (1, 5, 7, 7), (3, 3), (1, 1)),          # kernel as large as the image
    ((1, 1, 3, 70), (1, 1, 1, 3), (0, 2), (1, 2)),
]


@pytest.mark.parametrize("ishape,kshape,padding,strides", IM2COL_CASES)
def test_im2col_kernel(emu, ishape, kshape, padding, strides):
    B, C, H, W = ishape
    o = O.conv2d_out_shape(ishape, kshape, padding, strides)
    K, outHW = C * kshape[2] * kshape[3], o[2] * o[3]
    rng = np.random.default_rng(11)
    inp = rng.random(ishape, dtype=np.float32) + 1.0       # strictly positive: zero means padding
    geom = (ctypes.c_int64 * 11)(C, H, W, kshape[2], kshape[3], padding[0], padding[1], strides[0], strides[1], o[2], o[3])
    exp = np.stack([O.im2col(inp[b], ishape, kshape, padding, strides) for b in range(B)])
    for force_scalar in (0, 1):
        ws = np.full(B * K * outHW + 8, -5.0, np.float32)   # 8 guard elements
        v = emu.emu_im2col(ptr(ws), ptr(inp), B, geom, force_scalar)
        assert v == (4 if (outHW % 4 == 0 and not force_scalar) else 1)
        assert np.array_equal(ws[:B * K * outHW].reshape(B, K, outHW), exp)
        assert np.all(ws[B * K * outHW:] == -5.0)


def as_strided_idx(shape, strides, offset):
    idx = np.full(shape, offset, dtype=np.int64)
    for d, (n, s) in enumerate(zip(shape, strides)):
        sh = [1] * len(shape); sh[d] = n
        idx = idx + (np.arange(n, dtype=np.int64) * s).reshape(sh)
    return idx


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.uint16])
@pytest.mark.parametrize("shape,dst_strides,src_strides,merged", [
    ((6, 10), (10, 1), (1, 6), 2),                 # transposed source
    ((4, 5, 6), (30, 6, 1), (60, 12, 2), 1),       # every other element of the source: all dims merge
    ((3, 1, 7), (7, 7, 1), (14, 99, 2), 1),        # extent-1 dimension dropped, the rest merges
    ((2, 3, 4, 5), (60, 20, 5, 1), (1, 2, 6, 24), 4),
    ((8, 16), (16, 1), (32, 1), 2),                # padded rows
    ((5, 4, 3, 2, 2, 2), (96, 24, 8, 4, 2, 1), (1, 5, 20, 60, 120, 240), 6),
])
def test_copy_strided_kernel(emu, dt, shape, dst_strides, src_strides, merged):
    n_src = 1 + sum((n - 1) * abs(s) for n, s in zip(shape, src_strides))
    n_dst = 1 + sum((n - 1) * abs(s) for n, s in zip(shape, dst_strides))
    src = (np.arange(n_src) % 60000 + 1).astype(dt)
    dst = np.zeros(n_dst, dt)
    i64 = ctypes.c_int64
    arr = lambda t: (i64 * len(t))(*t)
    rank = emu.emu_copy_strided(np.dtype(dt).itemsize, ptr(dst), ptr(src), len(shape), arr(shape), arr(dst_strides),
                                arr(src_strides), 3)
    assert rank == merged
    exp = np.zeros(n_dst, dt)
    exp[as_strided_idx(shape, dst_strides, 0)] = src[as_strided_idx(shape, src_strides, 0)]
    assert np.array_equal(dst, exp)


FOREACH = {0: lambda x, y, z, a: x, 1: lambda x, y, z, a: np.full_like(x, a), 2: lambda x, y, z, a: a * x,
           3: lambda x, y, z, a: x + y, 4: lambda x, y, z, a: x - y, 5: lambda x, y, z, a: x * y,
           6: lambda x, y, z, a: x + y * z, 7: lambda x, y, z, a: a * x + y, 8: lambda x, y, z, a: x + y - np.sin(z)}


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("op", sorted(FOREACH))
@pytest.mark.parametrize("shape,layouts", [
    ((100, 70), ("c", "c", "c", "c")),                       # all contiguous: one merged dimension
    ((100, 70), ("c", "t", "c", "t")),                       # the reference's non-contiguous bench case (transposed inputs)
    ((6, 5, 4), ("s2", "c", "t", "s2")),                     # sliced output
])
def test_foreach_kernel(emu, dt, op, shape, layouts):
    rng = np.random.default_rng(op)
    n = int(np.prod(shape))

    def make(kind):
        if kind == "c":
            st = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]; size = n
        elif kind == "t":                                  # dimensions stored in reverse order
            st = [int(np.prod(shape[:i])) for i in range(len(shape))]; size = n
        else:                                              # every other element of a contiguous parent
            st = [2 * int(np.prod(shape[i + 1:])) for i in range(len(shape))]; size = 2 * n
        return rng.standard_normal(size).astype(dt), st
    bufs = [make(k) for k in layouts]
    idx = [as_strided_idx(shape, st, 0) for _, st in bufs]
    o0 = bufs[0][0].copy()
    alpha = 0.75
    want = FOREACH[op](bufs[1][0][idx[1]], bufs[2][0][idx[2]], bufs[3][0][idx[3]], dt(alpha))
    arr = lambda t: (i64 * len(t))(*t)
    rank = emu.emu_foreach(np.dtype(dt).itemsize, op, ptr(bufs[0][0]), ptr(bufs[1][0]), ptr(bufs[2][0]), ptr(bufs[3][0]),
                           len(shape), arr(shape), arr(bufs[0][1]), arr(bufs[1][1]), arr(bufs[2][1]), arr(bufs[3][1]), alpha, 3)
    assert rank == (1 if set(layouts) == {"c"} else len(shape))
    got = bufs[0][0]
    tol = 0 if op != 8 else (1e-6 if dt == np.float32 else 1e-15)
    assert np.abs(got[idx[0]] - want).max() <= tol * 4
    mask = np.ones(got.size, bool); mask[idx[0].reshape(-1)] = False
    assert np.array_equal(got[mask], o0[mask])              # only the elements the output view exposes are written


def test_foreach_in_place_update(emu):
    # `forEach x in a, y in b, z in c: x += y * z` (foreach.nim:231-232): the output aliases the first input
    shape = (33, 17)
    rng = np.random.default_rng(9)
    x = rng.standard_normal(shape).astype(np.float32); y = rng.standard_normal(shape).astype(np.float32)
    z = rng.standard_normal(shape).astype(np.float32)
    want = x + y * z
    arr = lambda t: (i64 * len(t))(*t)
    st = (17, 1)
    emu.emu_foreach(4, 6, ptr(x), ptr(x), ptr(y), ptr(z), 2, arr(shape), arr(st), arr(st), arr(st), arr(st), 0.0, 2)
    assert np.array_equal(x, want)
