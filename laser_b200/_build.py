"""Builds laser_b200/lib/liblaser_b200.so with nvcc for sm_100a (in-tree, no JIT cache)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liblaser_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["ptx.cuh", "f16_scale.cuh", "gemm_tc.cuh", "gemm_tc_kernel.inc", "gemm_simt.cuh", "gemm_simt_kernel.inc", "split.cuh", "layers.cuh", "capi_layers.inc",
           "../../include/laser_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",   # NOT -arch=sm_100a: tcgen05 needs the 'a' PTX target
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",                            # no libcuda/libcudart link dependency: loads on CPU-only hosts
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liblaser_b200.so")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source into the in-tree shared library. Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
