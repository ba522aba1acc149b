#!/bin/bash
# first GPU contact: smoke, staged parity tests, rounding probe + raw timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/host.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -15
echo "=== pytest golden+simt"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or simt or fill or degenerate" 2>&1 | tee gpurun_out/pytest_a.log | tail -15
echo "=== pytest rest"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not golden and not simt and not fill and not degenerate" 2>&1 | tee gpurun_out/pytest_b.log | tail -40
echo "=== perf probe"; timeout 600 python tools/perf_probe.py 2>&1 | tee gpurun_out/perf_probe.log | tail -40
