"""Host mirror of the reference's pre-packed GEMM API (gemm_prepacked.nim:63-292):

    gemm_prepackA_mem_required / gemm_prepackB_mem_required      -> bytes of device memory
    gemm_prepackA(dst, M, N, K, A, rowStrideA, colStrideA)       (gemm_prepacked.nim:194-218)
    gemm_prepackB(dst, M, N, K, B, rowStrideB, colStrideB)       (gemm_prepacked.nim:111-135)
    gemm_packed(M, N, K, alpha, packedA, packedB, beta, C, rowStrideC, colStrideC)   (:275-292)

On the GPU "packing" is the operand preparation of the default fp32-faithful mode (tf32 hi part
+ bf16 cross-term parts, compact K-major), so repeated products with a fixed matrix skip the
split pre-pass.  Buffers are opaque device memory (float32 only; device pointers only)."""
from ._capi import check, lib
from .gemm import _current_stream, _resolve
from .tensor import Storage

__all__ = ["gemm_prepackA_mem_required", "gemm_prepackB_mem_required", "gemm_prepackA", "gemm_prepackB",
           "gemm_packed", "gemm_packedB", "alloc_packed"]


def gemm_prepackA_mem_required(M, N, K):
    return int(lib().laser_b200_gemm_prepackA_mem_required_f32(M, N, K))


def gemm_prepackB_mem_required(M, N, K):
    return int(lib().laser_b200_gemm_prepackB_mem_required_f32(M, N, K))


def alloc_packed(nbytes):
    """Device buffer for a packed operand (256-byte aligned; freed with the returned object)."""
    return Storage(nbytes)


def _addr(buf):
    if isinstance(buf, Storage):
        return buf.raw_buffer
    p, _, dev = _resolve(buf)
    if not dev:
        raise TypeError("packed buffers live in device memory")
    return p


def _dev_f32(x):
    p, t, d = _resolve(x)
    if not d or t != "f32":
        raise TypeError("pre-packed GEMM takes float32 device pointers")
    return p


def gemm_prepackA(dst_packedA, M, N, K, A, rowStrideA, colStrideA, stream=None):
    stream = _current_stream() if stream is None else stream
    check(lib().laser_b200_gemm_prepackA_f32_dev(_addr(dst_packedA), M, N, K, _dev_f32(A), rowStrideA, colStrideA, stream))


def gemm_prepackB(dst_packedB, M, N, K, B, rowStrideB, colStrideB, stream=None):
    stream = _current_stream() if stream is None else stream
    check(lib().laser_b200_gemm_prepackB_f32_dev(_addr(dst_packedB), M, N, K, _dev_f32(B), rowStrideB, colStrideB, stream))


def gemm_packed(M, N, K, alpha, packedA, packedB, beta, C, rowStrideC, colStrideC, stream=None):
    stream = _current_stream() if stream is None else stream
    check(lib().laser_b200_gemm_packed_f32_dev(M, N, K, float(alpha), _addr(packedA), _addr(packedB), float(beta),
                                               _dev_f32(C), rowStrideC, colStrideC, stream))


def gemm_packedB(M, N, K, alpha, A, rowStrideA, colStrideA, packedB, beta, C, rowStrideC, colStrideC, stream=None):
    """A given as a plain strided matrix, B pre-packed (fixed weights)."""
    stream = _current_stream() if stream is None else stream
    check(lib().laser_b200_gemm_packedB_f32_dev(M, N, K, float(alpha), _dev_f32(A), rowStrideA, colStrideA,
                                                _addr(packedB), float(beta), _dev_f32(C), rowStrideC, colStrideC, stream))
