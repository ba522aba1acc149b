"""CPU-only: the Python mirror of the layer entry points (laser_b200/layers.py, tensor.py) and the
layer tests themselves (tests/test_gpu_zlayers.py) run against a CPU stand-in of the C-ABI library
(tests/emu/capi_python_emu.cpp: capi_layers.inc compiled for the host, kernels on host threads, every
GEMM through the emulated exact kernel).  This checks the ctypes marshalling, the view handling and
the tests' own expectations without a GPU; sizes that would take too long on host threads are skipped
there.  The stand-in is loaded only in the subprocess below (LASER_B200_LIB); the product never sees it."""
import os
import re
import subprocess
import sys

from emu_build import build_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_tests_pass_against_the_cpu_stand_in():
    so = build_emu("capi_python_emu", ["capi_layers.inc", "layers.cuh", "gemm_simt.cuh", "split.cuh",
                                        "../../include/laser_b200.h", "../../tests/emu/capi_layers_emu.cpp"])
    env = dict(os.environ, LASER_B200_LIB=so, LASER_B200_EMU="1", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_zlayers.py"), "-m", "gpu", "-q",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 50, tail
    assert "failed" not in out.stdout, tail
