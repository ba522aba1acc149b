#!/bin/bash
mkdir -p gpurun_out
for kc in 32 64 128; do echo "=== kc $kc"; LASER_B200_KC=$kc timeout 300 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/acc_kc$kc.log | grep -E "8192\^3|16384 P|8200 P"; done
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-600
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json | cut -c1-400
echo "=== ncu launch list (bench.py)"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/launches_bench.csv
echo "=== ncu full (tc kernels)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -o gpurun_out/prof_tc python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"split_rows|gemm_simt" -c 2 -o gpurun_out/prof_aux python tools/ncu_target.py > gpurun_out/ncu_aux.log 2>&1; tail -2 gpurun_out/ncu_aux.log
ls -la gpurun_out/
