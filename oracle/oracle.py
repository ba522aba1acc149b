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
]

ISA_NAMES = {1: "generic 2x1", 2: "avx+fma 6x16", 3: "avx512 14x32"}


def build(force=False):
    """Compile the oracle with its Makefile (gcc only; no GPU, no reference sources)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE])
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
