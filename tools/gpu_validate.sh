#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','clocks','e2e','gpu_launches','cpu_baseline'): print(k, d.get(k))
print({k:d['roofline'][k] for k in ('achieved','frac','tensor_pipe_frac','kernel_ms','prep_ms_per_step','traffic')}); print(d['modes'])"
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json | cut -c1-250
echo "=== ncu launch list (bench.py)"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench.csv
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 6 -o gpurun_out/prof_tc python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
echo "=== ncu full (split pre-pass)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:split_rows -c 2 -o gpurun_out/prof_split python tools/ncu_target.py > gpurun_out/ncu_split.log 2>&1; tail -1 gpurun_out/ncu_split.log
echo "=== layers (C++ self-check + HBM timings + sanitizer)"
LASER_B200_UNVALIDATED=1 timeout 900 python -m pytest tests/test_cpp_host.py -m gpu -q 2>&1 | tee gpurun_out/pytest_layers.log | tail -8
timeout 300 python tools/layers_bench.py > gpurun_out/layers_bench.jsonl 2> gpurun_out/layers_bench_err.log; cat gpurun_out/layers_bench.jsonl; tail -3 gpurun_out/layers_bench_err.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zlayers.py -m gpu -q -k "not 8192 and not 224" > gpurun_out/sanitizer_layers.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_layers.log | tail -3
echo "=== host-pointer entry: panel geometry variants"
for v in "" "LASER_B200_PANEL_ROWS=512" "LASER_B200_PANEL_TAPER=1" "LASER_B200_PANEL_ROWS=512 LASER_B200_PANEL_TAPER=1" "LASER_B200_PANEL_ROWS=2048 LASER_B200_PANEL_TAPER=1"; do env $v timeout 200 python tools/e2e_probe.py 2>> gpurun_out/e2e_probe_err.log | tee -a gpurun_out/e2e_probe.jsonl; done
echo "=== batched tensor-core launch (first run on a B200)"; LASER_B200_TC_BATCHED=1 timeout 600 python -m pytest tests/test_gpu_zlayers.py -m gpu -q -k "batched or conv2d" 2>&1 | tail -3; LASER_B200_TC_BATCHED=1 timeout 300 python tools/layers_bench.py 2>/dev/null | grep conv2d | sed "s/^/TC_BATCHED /"
echo "=== ncu full (layer kernels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"transpose_batched|im2col_kernel|foreach_strided" -c 8 -o gpurun_out/prof_layers python tools/ncu_layers_target.py > gpurun_out/ncu_layers.log 2>&1; tail -1 gpurun_out/ncu_layers.log
echo "=== L2 hints / raster (gemm_tc_hint_kernel next to the measured kernel)"
for v in "" "LASER_B200_L2HINT=1" "LASER_B200_L2HINT=2" "LASER_B200_L2HINT=1 LASER_B200_RASTER=6" "LASER_B200_L2HINT=1 LASER_B200_RASTER=4"; do echo "$v"; env $v timeout 300 python tools/perf_probe.py 2>> gpurun_out/perf_hint_err.log | tail -4 | tee -a gpurun_out/perf_hint.log; done
