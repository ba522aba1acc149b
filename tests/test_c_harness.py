"""The C ABI is usable from plain C: include/laser_b200.h compiles as C11 with -Wall -Werror,
links against the in-tree library (CPU), and a C caller reproduces the reference's known
answers on the GPU (the stand-in for the Nim call site, see tests/c_harness/gemm_harness.c)."""
import os
import subprocess

import pytest

import laser_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_harness", "gemm_harness.c")


def build(tmp_path):
    exe = str(tmp_path / "gemm_harness")
    libdir = os.path.dirname(L.lib_path())
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                           "-o", exe, "-L", libdir, "-llaser_b200", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_header_is_valid_c_and_links(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe, "--link-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "version" in out.stdout


@pytest.mark.gpu
def test_c_caller_on_gpu(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "max relative error" in out.stdout
