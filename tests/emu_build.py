"""Builds the host-thread emulation harnesses under tests/emu/ (TEST INFRASTRUCTURE, see cuda_emu.h)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "laser_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"


def build_emu(name, product_headers):
    """g++-compile tests/emu/<name>.cpp (which includes product kernel headers) into a shared library."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    out_dir = os.path.join(EMU_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "lib%s.so" % name)
    srcs = [os.path.join(EMU_DIR, name + ".cpp"), os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "ptx_emu.h")] + \
           [os.path.join(CSRC, h) for h in product_headers]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-pthread", "-fPIC", "-shared", "-ffp-contract=off",
                               "-I", CUDA_INC, "-I", EMU_DIR, "-Wno-attributes", "-Wno-unknown-pragmas", "-Wno-psabi", "-Wl,-Bsymbolic"   # stand-ins of CUDA runtime calls must win over a loaded libcudart
                              , srcs[0], "-o", so], env=env)
    return so
