#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke (pair)"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
rc=$?; echo "smoke rc=$rc"
echo "=== probe pair=1"; timeout 300 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/acc_pair1.log | grep -E "8192\^3|16384 P|8200 S"
echo "=== probe pair=0"; LASER_B200_CTA_PAIR=0 timeout 300 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/acc_pair0.log | grep -E "8192\^3"
echo "=== pytest gpu (pair)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -12
echo "=== perf probe"; timeout 600 python tools/perf_probe.py 2>&1 | tee gpurun_out/perf_probe.log | grep -E "n=8192|n=4096"
