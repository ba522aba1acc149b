"""CPU-only: the layer tests themselves (tests/test_gpu_zlayers.py) and with them the Python mirror of the
layer entry points (laser_b200/layers.py, tensor.py) run against the host-emulated build of the whole
library (tests/emu_build.py: build_capi_host_emu -- capi.cu compiled by g++ over stand-ins for the CUDA
runtime, kernels on host threads).  Sizes that would take too long on host threads are skipped there.
The emulated build is loaded only in the subprocess below (LASER_B200_LIB); the product never sees it."""
import os
import re
import subprocess
import sys

from emu_build import build_capi_host_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_tests_pass_against_the_host_emulated_library():
    so = build_capi_host_emu()
    env = dict(os.environ, LASER_B200_LIB=so, LASER_B200_EMU="1", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_zlayers.py"), "-m", "gpu", "-q",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 50, tail
    assert "failed" not in out.stdout, tail
