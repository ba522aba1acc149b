#!/bin/bash
# First B200 run of the two opt-in two-piece fp32 modes (bf16x3, f16x3): parity file, timing + error next to the default
# mode, launch list and ncu --set full of their kernels, compute-sanitizer.  gpurun --timeout 1500 -- 'bash tools/gpu_two_piece.sh'
mkdir -p gpurun_out
echo "=== pytest (two-piece modes)"; timeout 600 python -m pytest tests/test_gpu_zy_two_piece_modes.py -m gpu -q 2>&1 | tee gpurun_out/pytest_two_piece.log | tail -5
for m in bf16x3 f16x3; do
  echo "=== probe $m"; timeout 200 python tools/two_piece_probe.py $m 8192 10 2> gpurun_out/probe_${m}_err.log | tee gpurun_out/probe_${m}.json | cut -c1-1200
  echo "=== probe $m 4096"; timeout 200 python tools/two_piece_probe.py $m 4096 20 2>> gpurun_out/probe_${m}_err.log | tee gpurun_out/probe_${m}_4096.json | cut -c1-600
done
echo "=== accuracy of every fp32 mode vs the oracle and an fp64 product"; timeout 400 python tools/accuracy_probe.py 2>&1 | tee gpurun_out/accuracy_all_modes.log | cut -c1-400
echo "=== e2e in each mode (host pointers, pipelined)"
for m in tf32_bf16c bf16x3 f16x3; do echo $m; LASER_B200_F32_MODE=$m timeout 200 python tools/e2e_probe.py 2>> gpurun_out/e2e_modes_err.log | tee -a gpurun_out/e2e_modes.jsonl; done
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_piece.csv python tools/ncu_two_piece_target.py > gpurun_out/ncu_two_piece_list.log 2>&1; grep -c . gpurun_out/launches_two_piece.csv
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_f16_kernel|gemm_tc_kernel|absmax_mn|split_rows_f16x2|split_rows_bf16x2" -c 16 -o gpurun_out/prof_two_piece python tools/ncu_two_piece_target.py > gpurun_out/ncu_two_piece.log 2>&1; tail -1 gpurun_out/ncu_two_piece.log
echo "=== compute-sanitizer"; NCU_N=1024 timeout 600 compute-sanitizer --tool memcheck python tools/ncu_two_piece_target.py > gpurun_out/sanitizer_two_piece.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/sanitizer_two_piece.log
NCU_N=1024 timeout 600 compute-sanitizer --tool racecheck python tools/ncu_two_piece_target.py > gpurun_out/racecheck_two_piece.log 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck_two_piece.log
