"""CPU-only: the kernels whose timings are quoted (bench.py, profiles/) must be the kernels in
the shipped library.  profiles/sass_fingerprint.json holds an md5 of the SASS of every tensor-core GEMM
variant, of the exact kernel and of the operand-preparation / reduce / skinny-GEMM kernels as they were when last MEASURED on a
B200; this test recomputes them from the in-tree build (cuobjdump).  If a kernel was changed on purpose:
measure it on the GPU, then refresh the file with `python tests/test_sass_fingerprint.py --update`."""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILE = os.path.join(ROOT, "profiles", "sass_fingerprint.json")
MEASURED = ("gemm_tc_kernel", "gemm_simt_kernel", "split_rows_", "pack_general_kernel", "splitk_reduce_kernel", "gemv_warp", "fill_uniform_f32_kernel")


def fingerprints():
    sys.path.insert(0, ROOT)
    from laser_b200 import _build
    lib = _build.build()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, check=True).stdout
    fp, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1) if any(k in m.group(1) for k in MEASURED) else None
            if cur:
                fp[cur] = hashlib.md5()
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            fp[cur].update(line.split("*/", 1)[1].strip().encode())
    return {k: v.hexdigest() for k, v in sorted(fp.items())}


def test_measured_kernels_are_the_shipped_kernels():
    import pytest
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not installed")
    with open(FILE) as f:
        want = json.load(f)["kernels"]
    got = fingerprints()
    changed = sorted(k for k in want if got.get(k) != want[k])
    assert not changed, "SASS differs from the last measured build for: %s" % ", ".join(c[9:70] for c in changed)
    assert len(got) >= len(want)


if __name__ == "__main__":
    if "--update" in sys.argv:
        with open(FILE, "w") as f:
            json.dump({"note": "md5 of the SASS (cuobjdump -sass, encodings included) per kernel, as last measured on a B200",
                       "kernels": fingerprints()}, f, indent=1)
        print("wrote", FILE)
