#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -6
for kc in 128 256 512 1024 100000; do echo "=== kc $kc"; LASER_B200_KC=$kc timeout 300 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/acc_kc$kc.log | tail -12; done
echo "=== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
