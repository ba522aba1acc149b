"""CPU-only: the shipped library really is a tcgen05 / TMEM / TMA build (the SASS mnemonics of B200_PROFILING.md's
"what proves a Blackwell-native kernel" table), for every tensor-core kernel family, and holds no legacy tensor path.
`python tests/test_sass_evidence.py` rewrites the excerpt committed under profiles/."""
import collections
import os
import re
import shutil
import subprocess

import pytest

import laser_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "DMMA", "HMMA", "HGMMA", "ATOMG", "SYNCS", "STG.E.128",
             "USETMAXREG")


import functools


@functools.lru_cache(maxsize=1)
def sass_counts():
    """{kernel symbol: Counter(mnemonic -> count)} of the tensor-core kernels in liblaser_b200.so (one disassembly per run)"""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not installed")
    out = subprocess.run([exe, "-sass", L.lib_path()], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            # HMMA must not match UTCHMMA
            if re.search(r"(?<![A-Z])" + re.escape(mn), line):
                res[cur][mn] += 1
    return res


def test_tensor_core_kernels_are_tcgen05_tmem_tma():
    counts = sass_counts()
    tc = {k: v for k, v in counts.items() if "gemm_tc_kernel" in k}
    assert len(tc) == 32, sorted(tc)            # 4 families x (A, B major-ness) x (single CTA, CTA pair)
    for k, c in tc.items():
        assert c["UTCHMMA"] >= 4 and c["LDTM"] >= 8 and c["UTMALDG"] >= 2 and c["UTCBAR"] >= 2, (k, dict(c))
        assert c["USETMAXREG"] == 2, k          # warp-specialised register split
    for k, c in counts.items():
        assert c["HMMA"] == 0 and c["HGMMA"] == 0, k       # no half-precision mma.sync / wgmma anywhere in the library


def test_fp64_tensor_cores_and_bulk_copies():
    counts = sass_counts()
    dmma = [c for k, c in counts.items() if "gemm_dmma_kernel" in k]
    assert len(dmma) == 1 and dmma[0]["DMMA"] >= 32 and dmma[0]["LDGSTS"] >= 16      # mma.sync.m8n8k4.f64, cp.async staging
    ring = [c for k, c in counts.items() if "f16x2_rows_ring_kernel" in k]
    assert len(ring) == 1 and ring[0]["UBLKCP"] >= 2 and ring[0]["SYNCS"] >= 2          # cp.async.bulk rows + mbarrier


if __name__ == "__main__":
    counts = sass_counts()
    with open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w") as f:
        f.write("# cuobjdump -sass laser_b200/lib/liblaser_b200.so: occurrences of the Blackwell mnemonics per tensor-core kernel\n")
        f.write("# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor load, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,\n# LDGSTS = cp.async, DMMA = mma.sync.f64)\n")
        for k in sorted(counts):
            if "gemm_tc_kernel" in k or "gemm_dmma_kernel" in k or "f16x2_rows_ring_kernel" in k:
                f.write("%s\n    %s\n" % (k, "  ".join("%s=%d" % (m, counts[k][m]) for m in MNEMONICS if counts[k][m])))
    print("wrote profiles/r02_sass_mnemonics.txt")
