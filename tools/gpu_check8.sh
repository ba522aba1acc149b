#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -8
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','clocks','e2e','gpu_launches','roofline','modes','cpu_baseline'): print(k, d.get(k))"
tail -3 gpurun_out/bench_err.log
echo "=== ncu launch list (bench.py)"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench.csv
echo "=== ncu full (tc kernels)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 6 -o gpurun_out/prof_tc python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"split_rows" -c 2 -o gpurun_out/prof_aux python tools/ncu_target.py > gpurun_out/ncu_aux.log 2>&1; tail -1 gpurun_out/ncu_aux.log
