"""Builds the host-thread emulation harnesses under tests/emu/ (TEST INFRASTRUCTURE, see cuda_emu.h).

All of them are compiled with -fsanitize=alignment (abort on the first hit): a vector access (float4, the 4-element
transposition vectors, 16-byte bf16 stores ...) through a pointer that is not aligned to its type is a fault on the GPU
but silently works on x86, so the emulation would otherwise miss a forgotten alignment guard."""
import contextlib
import fcntl
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "laser_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"


@contextlib.contextmanager
def _build_lock(so):
    """pytest-xdist workers may ask for the same library at once: one builds (into a temporary name, renamed when
    complete), the others wait for the lock and find it up to date."""
    with open(so + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def _compile(cmd, so, env=None):
    tmp = "%s.tmp%d" % (so, os.getpid())
    subprocess.check_call(cmd + ["-o", tmp], env=env)
    os.replace(tmp, so)


def build_emu(name, product_headers):
    """g++-compile tests/emu/<name>.cpp (which includes product kernel headers) into a shared library."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    tsan = os.environ.get("LASER_B200_EMU_TSAN", "0") == "1"   # ThreadSanitizer variant (run pytest under LD_PRELOAD=libtsan)
    out_dir = os.path.join(EMU_DIR, "_build", "tsan") if tsan else os.path.join(EMU_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "lib%s.so" % name)
    srcs = [os.path.join(EMU_DIR, name + ".cpp"), os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "ptx_emu.h")] + \
           [os.path.join(CSRC, h) for h in product_headers]
    with _build_lock(so):
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            san = ["-O1", "-g", "-fsanitize=thread"] if tsan else ["-O2", "-fsanitize=alignment", "-fno-sanitize-recover=alignment"]
            _compile([gxx] + san + ["-std=c++17", "-pthread", "-fPIC", "-shared", "-ffp-contract=off",
                                    "-I", CUDA_INC, "-I", EMU_DIR, "-Wno-attributes", "-Wno-unknown-pragmas", "-Wno-psabi",
                                    "-Wl,-Bsymbolic",   # stand-ins of CUDA runtime calls must win over a loaded libcudart
                                    srcs[0]], so, env=env)
    return so


def _rewrite_launches(src):
    """`kernel<<<grid, block, smem, stream>>>(args)` -> `emu_launch_kernel(kernel, grid, block, smem, stream, args)`.
    The kernel expression is the identifier (with template arguments) that precedes `<<<`."""
    out, i = [], 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            break
        # walk back over the kernel expression: identifier chars, '::' and one balanced <...> template list
        k = j
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kernel = src[k:j]
        e = src.index(">>>", j)
        cfg = src[j + 3:e]
        assert src[e + 3] == "(", src[e:e + 20]
        depth, m = 0, e + 3
        while True:
            if src[m] == "(":
                depth += 1
            elif src[m] == ")":
                depth -= 1
                if depth == 0:
                    break
            m += 1
        args = src[e + 4:m]
        out.append(src[i:k])
        out.append("emu_launch_kernel(%s, %s%s%s)" % (kernel, cfg, ", " if args.strip() else "", args))
        i = m + 1
    return "".join(out)


def build_capi_host_emu(asan=False):
    """The whole host side of the library (laser_b200/csrc/capi.cu + capi_layers.inc) compiled for the CPU:
    generated translation unit = tests/emu/capi_host_prelude.h + capi.cu with its kernel launches rewritten."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    out_dir = os.path.join(EMU_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblaser_b200_hostemu_asan.so" if asan else "liblaser_b200_hostemu.so")
    csrc = os.path.abspath(CSRC)
    units = ["capi.cu", "tc_f16x3.cu", "tc_tf32x3.cu", "tc_tf32x1.cu", "tc_bf16.cu"]   # laser_b200/_build.py: SOURCES
    deps = [os.path.join(csrc, f) for f in units + ["capi_layers.inc", "f16_scale.cuh", "gemm_tc.cuh", "tc_params.h", "tc_launch.h",
                                                    "tc_launch_impl.cuh", "gemm_simt.cuh", "gemm_simt_kernel.inc", "split.cuh",
                                                    "layers.cuh", "ptx.cuh", "capi_multi.inc"] if os.path.exists(os.path.join(csrc, f))] + \
           [os.path.join(EMU_DIR, f) for f in ("capi_host_prelude.h", "cuda_emu.h", "ptx_emu.h")] + [os.path.abspath(__file__)]
    with _build_lock(so):
        return _build_capi_host_emu_locked(so, deps, units, csrc, out_dir, gxx, asan)


def _build_capi_host_emu_locked(so, deps, units, csrc, out_dir, gxx, asan):
    if os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    # ONE generated translation unit: the prelude, then every source of the library with its kernel launches rewritten
    parts = []
    for u in units:
        src = open(os.path.join(csrc, u)).read()
        src = _rewrite_launches(src)
        src = src.replace('#include "../../include/laser_b200.h"', '#include "%s"' % os.path.join(csrc, "..", "..", "include", "laser_b200.h"))
        assert "<<<" not in src and "cudaLaunchKernelEx" not in src
        parts.append("// ---- %s\n%s" % (u, src))
    assert parts[0].count("emu_launch_kernel(") >= 10
    gen = os.path.join(out_dir, "capi_host_emu_asan.cpp" if asan else "capi_host_emu.cpp")
    with open(gen, "w") as f:
        f.write('// GENERATED by tests/emu_build.py from laser_b200/csrc/*.cu -- do not edit\n#include "capi_host_prelude.h"\n' + "\n".join(parts))
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    san = ["-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else ["-O2"]
    _compile([gxx] + san + ["-std=c++17", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-fsanitize=alignment",
                            "-fno-sanitize-recover=alignment", "-I", CUDA_INC, "-I", EMU_DIR, "-I", csrc, "-Wno-attributes",
                            "-Wno-unknown-pragmas", "-Wno-psabi", "-Wl,-Bsymbolic", gen, "-ldl"], so, env=env)
    return so


def build_fake_nccl():
    """tests/emu/fake_nccl.c -> a shared library the emulated build loads through LASER_B200_NCCL_LIB"""
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    out_dir = os.path.join(EMU_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src, so = os.path.join(EMU_DIR, "fake_nccl.c"), os.path.join(out_dir, "libfake_nccl.so")
    with _build_lock(so):
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            _compile([gcc, "-O1", "-fPIC", "-shared", src], so)
    return so


def asan_env():
    """environment of a subprocess that loads an AddressSanitizer build through ctypes"""
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    lib = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not lib or not os.path.exists(lib):
        pytest.skip("libasan not installed")
    return {"LD_PRELOAD": os.path.realpath(lib), "ASAN_OPTIONS": "detect_leaks=0:abort_on_error=1"}
