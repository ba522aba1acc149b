#!/bin/bash
# Round 2, GPU session 1: baseline of the f16x3 path before it becomes the default (kc sweep, layouts, ncu).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2s1_smi.txt
for kc in 128 256 512; do
  echo "=== f16x3 8192 KC=$kc"; LASER_B200_KC=$kc timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>gpurun_out/r2s1_probe_err.log | tee gpurun_out/r2s1_probe_kc$kc.json | cut -c1-1500
done
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | tee gpurun_out/r2s1_layouts.log
echo "=== ncu dram single pass"; timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:"gemm_tc_f16_kernel|absmax_mn|split_rows_f16x2" -c 10 --csv --log-file gpurun_out/r2s1_dram.csv python tools/r2_ncu_f16_target.py > gpurun_out/r2s1_dram.log 2>&1; tail -3 gpurun_out/r2s1_dram.csv | cut -c1-300
echo "=== ncu full"; NCU_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_f16_kernel" -s 1 -c 1 -o gpurun_out/r2s1_f16_full python tools/r2_ncu_f16_target.py > gpurun_out/r2s1_full.log 2>&1; tail -2 gpurun_out/r2s1_full.log
NCU_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"absmax_mn|split_rows_f16x2" -c 4 -o gpurun_out/r2s1_prep_full python tools/r2_ncu_f16_target.py > gpurun_out/r2s1_prep.log 2>&1; tail -2 gpurun_out/r2s1_prep.log
ls -la gpurun_out | tail -12
