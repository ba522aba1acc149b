"""CPU-only: the oracle of the steps either side of the GEMM (oracle/laser_layers.c) against the
reference's own convolution known-answer vectors (conv2d_common.nim:128-283), against numpy for the
transposes (swapaxes.nim:16-112), and im2col+GEMM against the direct convolution."""
import json
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def conv_cases():
    with open(os.path.join(HERE, "golden", "conv2d_known_answer.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", conv_cases(), ids=lambda c: c["src"])
def test_conv2d_known_answer(case):
    inp = np.array(case["input"], np.float32); ker = np.array(case["kernel"], np.float32)
    tgt = np.array(case["target"], np.float32)
    args = (case["ishape"], ker, case["kshape"], case["padding"], case["strides"])
    assert O.conv2d_out_shape(case["ishape"], case["kshape"], case["padding"], case["strides"]) == tgt.shape
    assert np.array_equal(O.conv2d_im2col(inp, *args), tgt)
    assert np.array_equal(O.conv2d_direct(inp, *args), tgt)


def test_im2col_matrix_layout():
    # [C*kH*kW, outH*outW]: row (c, krow, kcol), column (oh, ow) -- conv2d_im2col.nim:69-93
    C, H, W, kH, kW, pH, pW, sH, sW = 2, 5, 6, 3, 2, 1, 0, 2, 1
    img = np.arange(C * H * W, dtype=np.float32).reshape(C, H, W) + 1
    ishape, kshape = (1, C, H, W), (4, C, kH, kW)
    _, _, oH, oW = O.conv2d_out_shape(ishape, kshape, (pH, pW), (sH, sW))
    ws = O.im2col(img, ishape, kshape, (pH, pW), (sH, sW))
    assert ws.shape == (C * kH * kW, oH * oW) == (O.im2col_workspace_size(ishape, kshape, (pH, pW), (sH, sW)) // (oH * oW), oH * oW)
    pad = np.zeros((C, H + 2 * pH, W + 2 * pW), np.float32); pad[:, pH:pH + H, pW:pW + W] = img
    for c in range(C):
        for kr in range(kH):
            for kc in range(kW):
                exp = pad[c, kr:kr + sH * oH:sH, kc:kc + sW * oW:sW][:oH, :oW]
                assert np.array_equal(ws[(c * kH + kr) * kW + kc].reshape(oH, oW), exp)


@pytest.mark.parametrize("ishape,kshape,padding,strides", [
    ((2, 3, 9, 11), (4, 3, 3, 3), (0, 0), (1, 1)),
    ((3, 2, 8, 8), (5, 2, 3, 3), (1, 1), (2, 2)),
    ((1, 4, 7, 10), (3, 4, 1, 1), (0, 0), (1, 1)),      # 1x1: im2col skipped (conv2d_im2col.nim:121)
    ((2, 1, 12, 6), (2, 1, 5, 2), (2, 1), (3, 3)),
])
def test_im2col_conv_matches_direct(ishape, kshape, padding, strides):
    rng = np.random.default_rng(3)
    inp = rng.integers(-3, 4, size=ishape).astype(np.float32)
    ker = rng.integers(-2, 3, size=kshape).astype(np.float32)
    a = O.conv2d_im2col(inp, ishape, ker, kshape, padding, strides)
    b = O.conv2d_direct(inp, ishape, ker, kshape, padding, strides)
    assert np.array_equal(a, b)          # small integers: exact whatever the summation order


def test_out_shape_rejects_bad_strides():
    with pytest.raises(ValueError):
        O.conv2d_out_shape((1, 1, 4, 4), (1, 1, 3, 3), (0, 0), (4, 1))   # conv2d_common.nim:35-36


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.int64, np.uint16, np.uint8])
@pytest.mark.parametrize("N,NR,NC", [(1, 1, 1), (1, 33, 65), (3, 32, 32), (2, 100, 7), (4, 5, 129)])
def test_transposes(dt, N, NR, NC):
    src = (np.arange(N * NR * NC) % 251).astype(dt).reshape(N, NR, NC)
    assert np.array_equal(O.transpose2D_batched(src, N, NR, NC), src.transpose(0, 2, 1))
    if N == 1:
        assert np.array_equal(O.transpose2D_copy(src[0], NR, NC), src[0].T)


def test_nchw_nhwc_round_trip():
    N, C, H, W = 2, 3, 5, 7
    x = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    y = O.nchw2nhwc(x, N, C, H, W)
    assert np.array_equal(y, x.transpose(0, 2, 3, 1))
    assert np.array_equal(O.nhwc2nchw(y, N, C, H, W), x)


def test_batched_gemm_is_a_loop_of_gemm_strided():
    batch, M, N, K = 3, 5, 7, 9
    rng = np.random.default_rng(0)
    A = rng.random((batch, M, K), dtype=np.float32); B = rng.random((K, N), dtype=np.float32)
    C = rng.random((batch, M, N), dtype=np.float32); C0 = C.copy()
    O.gemm_strided_batched(batch, M, N, K, 0.5, A, K, 1, M * K, B, N, 1, 0, 2.0, C, N, 1, M * N)   # shared B
    for b in range(batch):
        ref = C0[b].copy()
        O.gemm_strided(M, N, K, 0.5, A[b].copy(), K, 1, B, N, 1, 2.0, ref, N, 1)
        assert np.array_equal(C[b], ref)


def test_reference_1x1_shortcut_is_only_valid_for_unit_stride_without_padding():
    """conv2d_im2col.nim:121 skips im2col for every 1x1 kernel and reads the image in place; with a
    stride or padding that is a different (wrong) result than the reference's own direct convolution.
    The oracle restates the shortcut faithfully; the product goes through im2col for those cases
    (documented divergence, include/laser_b200.h) and is tested against im2col + GEMM instead."""
    ishape, kshape = (2, 4, 9, 9), (3, 4, 1, 1)
    rng = np.random.default_rng(0)
    inp = rng.integers(-3, 4, size=ishape).astype(np.float32); ker = rng.integers(-2, 3, size=kshape).astype(np.float32)
    assert np.array_equal(O.conv2d_im2col(inp, ishape, ker, kshape, (0, 0), (1, 1)), O.conv2d_direct(inp, ishape, ker, kshape, (0, 0), (1, 1)))
    direct = O.conv2d_direct(inp, ishape, ker, kshape, (1, 1), (2, 2))
    assert not np.array_equal(O.conv2d_im2col(inp, ishape, ker, kshape, (1, 1), (2, 2)), direct)
    M, K = kshape[0], ishape[1]
    o = O.conv2d_out_shape(ishape, kshape, (1, 1), (2, 2))
    via_im2col = np.stack([ker.reshape(M, K) @ O.im2col(inp[n], ishape, kshape, (1, 1), (2, 2)) for n in range(ishape[0])]).reshape(o)
    assert np.array_equal(via_im2col, direct)
