#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest prepacked"; timeout 900 python -m pytest tests/test_gpu_prepacked.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6
echo "=== perf"; timeout 600 python tools/perf_probe.py 2>&1 | tee gpurun_out/perf_probe.log | grep -E "prepack|n=8192"
