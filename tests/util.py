"""Shared helpers for the parity tests: strided layouts and golden vectors."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

LAYOUTS = ["row", "col", "rowslice", "colslice", "padded", "negrow", "negcol", "misaligned", "both2"]


def golden_cases():
    with open(os.path.join(HERE, "golden", "known_answer.json")) as f:
        return json.load(f)["cases"]


def embed(mat, layout):
    """Place the logical matrix `mat` (R x C) in a flat buffer with the given layout.
    Returns (flat buffer, offset of element (0,0), rowStride, colStride) in elements.
    Unused buffer slots are filled with a sentinel so that out-of-view reads/writes show."""
    R, C = mat.shape
    dt = mat.dtype
    if dt == np.uint16:
        sentinel = np.array(0xC2FA, dtype=dt)  # bf16 -125
    else:
        sentinel = np.array(-77, dtype=dt) if dt.kind in "iu" else np.array(-7777.0, dtype=dt)
    if layout == "row":
        buf = mat.reshape(-1).copy(); off, rs, cs = 0, C, 1
    elif layout == "col":
        buf = mat.T.reshape(-1).copy(); off, rs, cs = 0, 1, R
    elif layout == "rowslice":   # t[::2, :]
        big = np.full((2 * R, C), sentinel, dt); big[::2] = mat
        buf = big.reshape(-1); off, rs, cs = 0, 2 * C, 1
    elif layout == "colslice":   # t[:, ::2]
        big = np.full((R, 2 * C), sentinel, dt); big[:, ::2] = mat
        buf = big.reshape(-1); off, rs, cs = 0, 2 * C, 2
    elif layout == "padded":     # odd leading dimension
        big = np.full((R, C + 3), sentinel, dt); big[:, :C] = mat
        buf = big.reshape(-1); off, rs, cs = 0, C + 3, 1
    elif layout == "negrow":     # rows stored bottom-up
        buf = mat[::-1].reshape(-1).copy(); off, rs, cs = (R - 1) * C, -C, 1
    elif layout == "negcol":     # columns stored right-to-left
        buf = mat[:, ::-1].reshape(-1).copy(); off, rs, cs = C - 1, C, -1
    elif layout == "misaligned":  # contiguous but starting one element into the buffer
        buf = np.concatenate([np.full(1, sentinel, dt), mat.reshape(-1)]); off, rs, cs = 1, C, 1
    elif layout == "both2":      # t[::2, ::2] of a column-major parent
        big = np.full((2 * C, 2 * R), sentinel, dt); big[::2, ::2] = mat.T
        buf = big.reshape(-1); off, rs, cs = 0, 2, 4 * R
    else:
        raise ValueError(layout)
    return np.ascontiguousarray(buf), off, rs, cs


def extract(buf, off, rs, cs, R, C):
    """Read the logical R x C matrix back out of a flat buffer."""
    idx = off + np.arange(R)[:, None] * rs + np.arange(C)[None, :] * cs
    return buf[idx]


def f32_to_bf16_bits(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)
