#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -8
echo "=== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40
echo "=== perf probe"; timeout 600 python tools/perf_probe.py 2>&1 | tee gpurun_out/perf_probe.log | tail -40
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_err.log
