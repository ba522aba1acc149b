"""TEST INFRASTRUCTURE ONLY -- numpy-facing wrapper of oracle/liblaser_oracle.so.

Mirrors the reference signature
    gemm_strided(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB,
                 beta, C, rowStrideC, colStrideC)          (gemm.nim:184-193)
on numpy buffers.  See laser_oracle.h for what is restated and how parity is pinned.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblaser_oracle.so")
_lib = None

__all__ = [
    "build", "lib", "gemm_strided", "cpu_gemm_strided_f32", "gemm_f32_in_f64",
    "fill_uniform_f32", "mean_relative_error", "max_relative_error",
    "normwise_relative_error", "detect_isa", "num_threads", "set_num_threads", "ISA_NAMES",
    "transpose2D_copy", "transpose2D_batched", "nchw2nhwc", "nhwc2nchw", "conv2d_out_shape",
    "im2col_workspace_size", "im2col", "conv2d_im2col", "conv2d_direct", "gemm_strided_batched",
]

ISA_NAMES = {1: "generic 2x1", 2: "avx+fma 6x16", 3: "avx512 14x32"}


def build(force=False):
    """Compile the oracle with its Makefile (gcc only; no GPU, no reference sources)."""
    if force or not os.path.exists(_LIB_PATH):
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:      # one process builds, the others of a multi-rank job wait
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or not os.path.exists(_LIB_PATH):
                    subprocess.check_call(["make", "-s", "-C", _HERE])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, f32, f64, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p
        for name, sc in (("f32", f32), ("f64", f64), ("i32", ctypes.c_int32), ("i64", i64),
                         ("bf16", f32)):
            fn = getattr(L, "oracle_gemm_strided_" + name)
            fn.restype = None
            fn.argtypes = [i64, i64, i64, sc, vp, i64, i64, vp, i64, i64, sc, vp, i64, i64]
        L.laser_cpu_gemm_strided_f32.restype = ctypes.c_int
        L.laser_cpu_gemm_strided_f32.argtypes = [i64, i64, i64, f32, vp, i64, i64, vp, i64, i64,
                                                 f32, vp, i64, i64, ctypes.c_int]
        L.oracle_gemm_f32_in_f64.restype = None
        L.oracle_gemm_f32_in_f64.argtypes = [i64, i64, i64, vp, i64, i64, vp, i64, i64, vp]
        for n in ("mean", "max", "normwise"):
            fn = getattr(L, "oracle_%s_relative_error_f32" % n)
            fn.restype = f64
            fn.argtypes = [vp, vp, i64]
        L.oracle_fill_uniform_f32.restype = None
        L.oracle_fill_uniform_f32.argtypes = [vp, i64, ctypes.c_uint64, f32, f32]
        L.laser_cpu_detect_isa.restype = ctypes.c_int
        L.laser_cpu_num_threads.restype = ctypes.c_int
        L.laser_cpu_set_num_threads.restype = None
        L.laser_cpu_set_num_threads.argtypes = [ctypes.c_int]
        I4, I2 = i64 * 4, i64 * 2
        L.oracle_transpose2d_copy.restype = None
        L.oracle_transpose2d_copy.argtypes = [vp, vp, i64, i64, ctypes.c_int]
        L.oracle_transpose2d_batched.restype = None
        L.oracle_transpose2d_batched.argtypes = [vp, vp, i64, i64, i64, ctypes.c_int]
        for n in ("oracle_nchw2nhwc", "oracle_nhwc2nchw"):
            getattr(L, n).restype = None
            getattr(L, n).argtypes = [vp, vp, i64, i64, i64, i64, ctypes.c_int]
        L.oracle_conv2d_out_shape.restype = ctypes.c_int
        L.oracle_conv2d_out_shape.argtypes = [I4, I4, I2, I2, I4]
        L.oracle_im2col_workspace_size.restype = i64
        L.oracle_im2col_workspace_size.argtypes = [I4, I4, I2, I2]
        L.oracle_im2col_f32.restype = None
        L.oracle_im2col_f32.argtypes = [vp, i64, i64, vp] + [i64] * 9
        L.oracle_conv2d_im2col_f32.restype = ctypes.c_int
        L.oracle_conv2d_im2col_f32.argtypes = [vp, vp, I4, vp, I4, I2, I2, vp]
        L.oracle_conv2d_direct_f32.restype = ctypes.c_int
        L.oracle_conv2d_direct_f32.argtypes = [vp, vp, I4, vp, I4, I2, I2]
        L.oracle_gemm_strided_batched_f32.restype = None
        L.oracle_gemm_strided_batched_f32.argtypes = [i64, i64, i64, i64, f32, vp, i64, i64, i64, vp, i64, i64, i64,
                                                      f32, vp, i64, i64, i64]
        _lib = L
    return _lib


_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64",
           np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def gemm_strided(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, bf16=False):
    """Numerics-faithful oracle.  A, B, C are numpy arrays used as raw buffers
    (the pointer is element [0]); strides are in ELEMENTS, as in the reference.
    bf16=True: buffers are uint16 bf16 bit patterns."""
    if bf16:
        assert A.dtype == np.uint16 and B.dtype == np.uint16 and C.dtype == np.uint16
        fn = lib().oracle_gemm_strided_bf16
    else:
        assert A.dtype == B.dtype == C.dtype
        fn = getattr(lib(), "oracle_gemm_strided_" + _SUFFIX[A.dtype])
    fn(M, N, K, alpha, _ptr(A), rsA, csA, _ptr(B), rsB, csB, beta, _ptr(C), rsC, csC)
    return C


def cpu_gemm_strided_f32(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, isa=0):
    """Structure-faithful restatement (packing + micro-kernel + OpenMP). Returns isa used."""
    assert A.dtype == B.dtype == C.dtype == np.float32
    return lib().laser_cpu_gemm_strided_f32(M, N, K, alpha, _ptr(A), rsA, csA, _ptr(B), rsB, csB,
                                            beta, _ptr(C), rsC, csC, isa)


def gemm_f32_in_f64(M, N, K, A, rsA, csA, B, rsB, csB):
    C = np.empty((M, N), dtype=np.float64)
    lib().oracle_gemm_f32_in_f64(M, N, K, _ptr(A), rsA, csA, _ptr(B), rsB, csB, _ptr(C))
    return C


def fill_uniform_f32(n, seed, lo, hi):
    out = np.empty(int(n), dtype=np.float32)
    lib().oracle_fill_uniform_f32(_ptr(out), int(n), int(seed), lo, hi)
    return out


def _flat32(x):
    return np.ascontiguousarray(x, dtype=np.float32).reshape(-1)


def mean_relative_error(y, y_true):
    """laser/private/error_functions.nim:19-26 (symmetric denominator)."""
    y, t = _flat32(y), _flat32(y_true)
    return lib().oracle_mean_relative_error_f32(_ptr(y), _ptr(t), y.size)


def max_relative_error(y, y_true):
    """max |y - y_true| / |y_true| (BASELINE.json gate)."""
    y, t = _flat32(y), _flat32(y_true)
    return lib().oracle_max_relative_error_f32(_ptr(y), _ptr(t), y.size)


def normwise_relative_error(y, y_true):
    y, t = _flat32(y), _flat32(y_true)
    return lib().oracle_normwise_relative_error_f32(_ptr(y), _ptr(t), y.size)


def detect_isa():
    return lib().laser_cpu_detect_isa()


def num_threads():
    return lib().laser_cpu_num_threads()


def set_num_threads(n):
    lib().laser_cpu_set_num_threads(int(n))


# ---- the steps either side of the GEMM (laser_layers.c) --------------------------------------
def _i4(t):
    return (ctypes.c_int64 * 4)(*[int(v) for v in t])


def _i2(t):
    return (ctypes.c_int64 * 2)(*[int(v) for v in t])


def transpose2D_batched(src, N, NR, NC):
    """swapaxes.nim:56-81 on a contiguous buffer of N matrices [NR, NC]; returns [N, NC, NR]."""
    src = np.ascontiguousarray(src)
    dst = np.empty(N * NR * NC, dtype=src.dtype)
    lib().oracle_transpose2d_batched(_ptr(dst), _ptr(src), N, NR, NC, src.dtype.itemsize)
    return dst.reshape(N, NC, NR)


def transpose2D_copy(src, NR, NC):
    """swapaxes.nim:16-54."""
    src = np.ascontiguousarray(src)
    dst = np.empty(NR * NC, dtype=src.dtype)
    lib().oracle_transpose2d_copy(_ptr(dst), _ptr(src), NR, NC, src.dtype.itemsize)
    return dst.reshape(NC, NR)


def nchw2nhwc(src, N, C, H, W):
    src = np.ascontiguousarray(src)
    dst = np.empty(N * C * H * W, dtype=src.dtype)
    lib().oracle_nchw2nhwc(_ptr(dst), _ptr(src), N, C, H, W, src.dtype.itemsize)
    return dst.reshape(N, H, W, C)


def nhwc2nchw(src, N, C, H, W):
    src = np.ascontiguousarray(src)
    dst = np.empty(N * C * H * W, dtype=src.dtype)
    lib().oracle_nhwc2nchw(_ptr(dst), _ptr(src), N, C, H, W, src.dtype.itemsize)
    return dst.reshape(N, C, H, W)


def conv2d_out_shape(ishape, kshape, padding, strides):
    """conv2d_common.nim:15-45; (n, c, h, w) of the output."""
    out = (ctypes.c_int64 * 4)()
    if lib().oracle_conv2d_out_shape(_i4(ishape), _i4(kshape), _i2(padding), _i2(strides), out):
        raise ValueError("strides must satisfy 0 < s < extent (conv2d_common.nim:35-36)")
    return tuple(out)


def im2col_workspace_size(ishape, kshape, padding, strides):
    return int(lib().oracle_im2col_workspace_size(_i4(ishape), _i4(kshape), _i2(padding), _i2(strides)))


def im2col(image, ishape, kshape, padding, strides):
    """One image [C, H, W] -> [C*kH*kW, outH*outW] (conv2d_im2col.nim:44-93)."""
    o = conv2d_out_shape(ishape, kshape, padding, strides)
    image = np.ascontiguousarray(image, dtype=np.float32)
    K = ishape[1] * kshape[2] * kshape[3]
    ws = np.empty(K * o[2] * o[3], dtype=np.float32)
    lib().oracle_im2col_f32(_ptr(ws), o[2], o[3], _ptr(image), ishape[1], ishape[2], ishape[3], kshape[2],
                            kshape[3], padding[0], padding[1], strides[0], strides[1])
    return ws.reshape(K, o[2] * o[3])


def conv2d_im2col(inp, ishape, kernel, kshape, padding, strides):
    """conv2d_im2col.nim:95-166; NCHW in, NCHW out."""
    o = conv2d_out_shape(ishape, kshape, padding, strides)
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    out = np.zeros(int(np.prod(o)), dtype=np.float32)
    ws = np.empty(max(1, im2col_workspace_size(ishape, kshape, padding, strides)), dtype=np.float32)
    rc = lib().oracle_conv2d_im2col_f32(_ptr(out), _ptr(inp), _i4(ishape), _ptr(kernel), _i4(kshape), _i2(padding),
                                        _i2(strides), _ptr(ws))
    if rc:
        raise ValueError("oracle_conv2d_im2col_f32 -> %d" % rc)
    return out.reshape(o)


def conv2d_direct(inp, ishape, kernel, kshape, padding, strides):
    """conv2d_direct_convolution.nim:8-76 (cross-check)."""
    o = conv2d_out_shape(ishape, kshape, padding, strides)
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    out = np.zeros(int(np.prod(o)), dtype=np.float32)
    rc = lib().oracle_conv2d_direct_f32(_ptr(out), _ptr(inp), _i4(ishape), _ptr(kernel), _i4(kshape), _i2(padding),
                                        _i2(strides))
    if rc:
        raise ValueError("oracle_conv2d_direct_f32 -> %d" % rc)
    return out.reshape(o)


def gemm_strided_batched(batch, M, N, K, alpha, A, rsA, csA, bsA, B, rsB, csB, bsB, beta, C, rsC, csC, bsC):
    assert A.dtype == B.dtype == C.dtype == np.float32
    lib().oracle_gemm_strided_batched_f32(batch, M, N, K, alpha, _ptr(A), rsA, csA, bsA, _ptr(B), rsB, csB, bsB,
                                          beta, _ptr(C), rsC, csC, bsC)
    return C
