"""CPU-only: host-side decisions of the C library that need no GPU -- operand classification for
the TMA path, address spans of strided views (what the host-pointer entry
copies), pre-packed buffer sizing."""
import ctypes

import laser_b200 as L
from laser_b200._capi import lib

K_MAJOR, MN_MAJOR, GENERAL = 0, 1, 2


def classify(esz, base, s_mn, s_k):
    return lib().laser_b200_debug_classify(esz, ctypes.c_void_p(base), s_mn, s_k)


def test_operand_classification():
    a = 0x7f0000000000                              # 16-byte aligned base
    assert classify(4, a, 8192, 1) == K_MAJOR         # row-major A (rowStride = K, colStride = 1)
    assert classify(4, a, 1, 8192) == MN_MAJOR        # A given transposed / ordinary row-major B
    assert classify(4, a, 8191, 1) == GENERAL         # odd leading dimension: not a 16-byte multiple
    assert classify(4, a + 4, 8192, 1) == GENERAL     # misaligned base
    assert classify(4, a, 16384, 2) == GENERAL        # t[:, ::2]
    assert classify(4, a, -8192, 1) == GENERAL        # negative stride
    assert classify(4, a, 0, 1) == GENERAL
    assert classify(2, a, 8, 1) == K_MAJOR and classify(2, a, 4, 1) == GENERAL   # bf16: 8 elements = 16 bytes
    assert classify(4, a, 1, 1) == GENERAL            # both strides 1 (a vector): neither pitch is a 16-byte multiple
    assert classify(4, a, 2**40, 1) == GENERAL        # pitch beyond what a tensor map can encode


def span(rows, cols, rs, cs):
    lo, hi, dense = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    assert lib().laser_b200_debug_span(rows, cols, rs, cs, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(dense)) == 0
    return lo.value, hi.value, bool(dense.value)


def test_view_spans():
    assert span(3, 4, 4, 1) == (0, 11, True)          # contiguous row-major
    assert span(3, 4, 1, 3) == (0, 11, True)          # contiguous column-major
    assert span(3, 4, 8, 1) == (0, 19, False)         # padded rows: gaps must survive the round trip
    assert span(3, 4, -4, 1) == (-8, 3, True)         # rows stored bottom-up
    assert span(3, 4, 4, -1) == (-3, 8, True)         # columns right-to-left
    assert span(3, 4, 8, 2) == (0, 22, False)
    assert span(1, 1, 7, 9) == (0, 0, True)


def test_prepack_sizes():
    M, N, K = 1000, 640, 513
    need_a, need_b = L.gemm_prepackA_mem_required(M, N, K), L.gemm_prepackB_mem_required(M, N, K)
    # two fp16 arrays (pitch K rounded to 8) + one abs-max word per row, 256-byte aligned sections
    assert 2 * M * 520 * 2 + M * 4 <= need_a <= 2 * M * 520 * 2 + M * 4 + 3 * 256 and need_a % 256 == 0
    assert 2 * N * 520 * 2 + N * 4 <= need_b <= 2 * N * 520 * 2 + N * 4 + 3 * 256 and need_b % 256 == 0
    assert L.gemm_prepackA_mem_required(0, N, K) == 0


def test_library_is_current_by_content_not_by_modification_time(tmp_path):
    """The tree that travels to a GPU box is a COPY: contents survive, modification times do not.  The build records a
    digest of the sources next to the library; a touched source must not trigger a rebuild (8 ranks importing the
    package on such a box once raced 8 rebuilds), an edited one must."""
    import os
    from laser_b200 import _build as B
    B.build()
    assert B._stamp() == B._src_digest() and not B.needs_build()
    src = os.path.join(B.CSRC, "capi.cu")
    st = os.stat(src)
    try:
        os.utime(src)                                   # newer than the library now
        assert os.path.getmtime(src) > os.path.getmtime(B.LIB_PATH)
        assert not B.needs_build()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    saved = B._stamp()
    try:
        with open(B.STAMP_PATH, "w") as f:
            f.write("0" * 64 + "\n")                    # what an edited source looks like
        assert B.needs_build()
    finally:
        with open(B.STAMP_PATH, "w") as f:
            f.write(saved + "\n")
    assert not B.needs_build()
