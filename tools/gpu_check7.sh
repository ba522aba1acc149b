#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== probe"; timeout 300 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/acc_mixed.log | cut -c1-330
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
