#!/bin/bash
# Round 2, GPU session 11: A/B of two builds of the library in one call (LASER_B200_LIB): the epilogue that computes a
# 32-column chunk before waiting for the staging buffer vs the previous build.
mkdir -p gpurun_out
for i in 1 2 3; do
for v in "X=1" "LASER_B200_LIB=laser_b200/lib/prev/liblaser_b200_prev.so"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>gpurun_out/r2s11_err.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.3f kernel %.3f prep %.3f' % (d['ms'],d['kernel_ms'],d['prep_ms_per_step']))"; done; done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,sm__cycles_elapsed.max
for v in "X=1" "LASER_B200_LIB=laser_b200/lib/prev/liblaser_b200_prev.so"; do env $v NCU_REPS=2 timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel" -c 2 --csv --log-file gpurun_out/r2s11_m.csv python tools/r2_ncu_f16_target.py > /dev/null 2>&1; echo "--- $v"; grep gemm_tc gpurun_out/r2s11_m.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' '; echo; done
