"""CPU-only: the C-ABI library builds, loads without a GPU/driver, exports every symbol
include/laser_b200.h declares, and FAILS LOUDLY (no CPU fallback) when asked to compute
without a device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import laser_b200 as L
from laser_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "laser_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(laser_b200_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_capi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.lib_path()], text=True)
    exported = set(re.findall(r" T (laser_b200_\w+)", out))
    assert set(declared_symbols()) <= exported


def test_no_link_dependency_on_driver_or_torch():
    out = subprocess.check_output(["ldd", L.lib_path()], text=True)
    assert "libcuda.so" not in out and "torch" not in out and "libcudart" not in out


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "laser_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liblaser_oracle" not in src, f


def test_compute_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    a = np.ones((4, 4), np.float32); c = np.zeros((4, 4), np.float32)
    with pytest.raises(L.LaserB200Error) as ei:
        L.gemm_strided(4, 4, 4, 1.0, a, 4, 1, a, 4, 1, 0.0, c, 4, 1)
    assert ei.value.code == _capi.E_NODEVICE
    assert np.all(c == 0)  # nothing computed behind our back


def test_tensor_contract_host_side():
    """shape/strides/offset bookkeeping of the Tensor mirror (no device memory needed)."""
    from laser_b200.tensor import Storage, Tensor, _row_major_strides
    assert _row_major_strides([2, 3, 4]) == [12, 4, 1]       # initialization.nim:24-32
    st = Storage(nbytes=4 * 24, ptr=0x1000, owner=False)
    t = Tensor([4, 6], [6, 1], 0, st, "f32")
    assert t.rank == 2 and t.size == 24 and t.is_C_contiguous()
    tt = t.transpose()
    assert tt.shape == [6, 4] and tt.strides == [1, 6] and not tt.is_C_contiguous()
    s = t.slice2d(slice(1, 4, 2), slice(2, 6))
    assert s.shape == [2, 4] and s.strides == [12, 1] and s.offset == 8
    assert s.unsafe_raw_data() == 0x1000 + 8 * 4                # datatypes.nim:64-88
    assert Tensor([1, 5], [99, 1], 0, st, "f32").is_C_contiguous()  # size-1 dims ignore strides
    with pytest.raises(ValueError):
        Tensor([1] * 7, [1] * 7, 0, st, "f32")                   # LASER_MAXRANK = 6
