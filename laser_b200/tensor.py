"""Device Tensor honouring the reference's tensor contract.

Mirrors laser/tensor/datatypes.nim:12-88 and initialization.nim:24-202:
  Tensor{shape, strides, offset, storage}   (strides/offset in ELEMENTS, rank <= 6,
                                             row-major by default)
  rank, size, is_C_contiguous, unsafe_raw_data (= address of element [0,...,0] incl. offset)
  newTensor(shape) (zero-initialised, like setZero), toTensor(data), copyFrom, deepCopy
Storage lives in HBM and is owned by a reference-counted Storage object (CpuStorage is a
`ref object` with memowner, datatypes.nim:24-30); slicing/transposing share it.
"""
import ctypes

import numpy as np

from ._capi import MAXRANK, TensorView, check, lib, vp

__all__ = ["Tensor", "Storage", "newTensor", "toTensor", "matmul", "LASER_MAXRANK"]

LASER_MAXRANK = MAXRANK  # laser/dynamic_stack_arrays.nim:6
_ITEMSIZE = {"f32": 4, "f64": 8, "i32": 4, "i64": 8, "bf16": 2}
_CODE = {"f32": 0, "f64": 1, "i32": 2, "i64": 3, "bf16": 4}
_NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64, "bf16": np.uint16}
_FROM_NP = {np.dtype(v): k for k, v in _NP.items()}


class Storage:
    """Device analogue of CpuStorage (datatypes.nim:24-30): raw_buffer + ownership flag."""

    def __init__(self, nbytes=0, ptr=None, owner=True, keepalive=None):
        self.nbytes = int(nbytes)
        self.memowner = owner
        self._keepalive = keepalive
        if ptr is None:
            p = vp()
            check(lib().laser_b200_malloc(ctypes.byref(p), self.nbytes))
            self.raw_buffer = int(p.value)
        else:
            self.raw_buffer = int(ptr)

    def __del__(self):
        try:
            if self.memowner and self.raw_buffer:
                lib().laser_b200_free(self.raw_buffer)
                self.raw_buffer = 0
        except Exception:
            pass


def _row_major_strides(shape):
    strides, acc = [], 1
    for d in reversed(shape):
        strides.append(acc)
        acc *= int(d)
    return list(reversed(strides))


class Tensor:
    def __init__(self, shape, strides, offset, storage, dtype):
        if len(shape) > LASER_MAXRANK:
            raise ValueError("rank %d > LASER_MAXRANK=%d" % (len(shape), LASER_MAXRANK))
        self.shape = [int(s) for s in shape]
        self.strides = [int(s) for s in strides]
        self.offset = int(offset)
        self.storage = storage
        self.dtype = dtype

    # ---- datatypes.nim:32-47 -------------------------------------------------
    @property
    def rank(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def is_C_contiguous(self):
        cur = 1
        for i in range(self.rank - 1, -1, -1):
            if self.shape[i] != 1 and self.strides[i] != cur:
                return False
            cur *= self.shape[i]
        return True

    # ---- datatypes.nim:64-88 -------------------------------------------------
    def unsafe_raw_data(self):
        """Device address of element [0, ..., 0] (storage + offset), as an int."""
        return self.storage.raw_buffer + self.offset * _ITEMSIZE[self.dtype]

    # ---- views (share storage, like the reference's shallow CpuStorage) ------
    def transpose(self):
        if self.rank != 2:
            raise ValueError("transpose needs rank 2")
        return Tensor(self.shape[::-1], self.strides[::-1], self.offset, self.storage, self.dtype)

    def slice2d(self, rows=slice(None), cols=slice(None)):
        if self.rank != 2:
            raise ValueError("slice2d needs rank 2")
        shape, strides, off = [], [], self.offset
        for dim, sl in enumerate((rows, cols)):
            start, stop, step = sl.indices(self.shape[dim])
            n = len(range(start, stop, step))
            off += start * self.strides[dim]
            shape.append(n)
            strides.append(self.strides[dim] * step)
        return Tensor(shape, strides, off, self.storage, self.dtype)

    # ---- host <-> device ------------------------------------------------------
    def copyFrom(self, data):
        """copyFromRaw analogue (initialization.nim:80-128) for a C-contiguous tensor."""
        a = np.ascontiguousarray(data, dtype=_NP[self.dtype])
        if not self.is_C_contiguous() or a.size != self.size:
            raise ValueError("copyFrom needs a C-contiguous tensor of matching size")
        check(lib().laser_b200_memcpy_h2d(self.unsafe_raw_data(), a.ctypes.data, a.nbytes))
        return self

    def to_numpy(self):
        """Copies the whole storage back and re-applies shape/strides/offset on the host."""
        itemsize = _ITEMSIZE[self.dtype]
        n = self.storage.nbytes // itemsize
        host = np.empty(n, dtype=_NP[self.dtype])
        check(lib().laser_b200_memcpy_d2h(host.ctypes.data, self.storage.raw_buffer, n * itemsize))
        return np.lib.stride_tricks.as_strided(
            host[self.offset:], shape=self.shape, strides=[s * itemsize for s in self.strides]).copy()

    def deepCopy(self):
        out = newTensor(self.shape, self.dtype)
        out.copyFrom(self.to_numpy())
        return out

    def view_struct(self):
        v = TensorView()
        v.rank = self.rank
        v.dtype = _CODE[self.dtype]
        for i in range(self.rank):
            v.shape[i] = self.shape[i]
            v.strides[i] = self.strides[i]
        v.offset = self.offset
        v.storage = self.storage.raw_buffer
        return v

    @staticmethod
    def from_torch(t):
        """Non-owning view of a torch CUDA tensor (keeps it alive)."""
        from .gemm import _torch_dtype_name
        name = _torch_dtype_name(t)
        if name is None or not t.is_cuda:
            raise TypeError("need a CUDA tensor of a supported dtype")
        st = Storage(nbytes=t.untyped_storage().nbytes(), ptr=t.untyped_storage().data_ptr(),
                     owner=False, keepalive=t)
        return Tensor(list(t.shape), list(t.stride()), t.storage_offset(), st, name)


def newTensor(shape, dtype="f32"):
    """Zero-initialised row-major tensor (initialization.nim:156-170)."""
    shape = [int(s) for s in shape]
    n = 1
    for s in shape:
        n *= s
    st = Storage(n * _ITEMSIZE[dtype])
    check(lib().laser_b200_memset_zero(st.raw_buffer, st.nbytes))
    return Tensor(shape, _row_major_strides(shape), 0, st, dtype)


def toTensor(data, dtype=None):
    """Host array -> device tensor (initialization.nim:172-202)."""
    a = np.asarray(data)
    if dtype is None:
        dtype = _FROM_NP.get(a.dtype)
        if dtype is None:
            raise TypeError("unsupported dtype %s" % a.dtype)
    t = newTensor(a.shape, dtype)
    return t.copyFrom(a)


def matmul(A, B, C=None, alpha=1.0, beta=0.0, path=0, stream=None):
    """C <- alpha*A@B + beta*C on rank-2 device tensors of any strides: the tensor-level
    caller of gemm_strided (gemm_prepacked.nim:306-307 passes t.unsafe_raw_data)."""
    from .gemm import _current_stream
    if C is None:
        C = newTensor([A.shape[0], B.shape[1]], A.dtype)
    va, vb, vc = A.view_struct(), B.view_struct(), C.view_struct()
    if stream is None:
        stream = _current_stream()
    check(lib().laser_b200_matmul_views(ctypes.byref(va), ctypes.byref(vb), ctypes.byref(vc),
                                        float(alpha), float(beta), int(path), stream))
    return C
