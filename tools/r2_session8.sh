#!/bin/bash
# Round 2, GPU session 8: ticketed one-launch preparation of MN-major operands (second read from L2), few-rows kernel of the
# im2col convolution, fp64 DMMA kernel with cp.async staging; full GPU suite.
mkdir -p gpurun_out
O=gpurun_out
echo "=== probes"
for v in "X=1" "LASER_B200_PREP_TICKET=0" "X=2" "LASER_B200_PREP_TICKET=0"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>$O/r2s8_err.log | tee -a $O/r2s8_probes.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.3f kernel %.3f prep %.3f mre %.2e' % (d['ms'],d['kernel_ms'],d['prep_ms_per_step'],d['error_vs_fp64_S_U(-0.1,0.1)']['f16x3']['mean_relative_error']))"; done
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | grep -E "f16x3|M=32768" | tee $O/r2s8_layouts.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2s8_pytest.log
echo "=== ncu metrics (single pass)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows|f16x2_cols" -c 6 --csv --log-file $O/r2s8_metrics.csv python tools/r2_ncu_f16_target.py > $O/r2s8_metrics.log 2>&1; grep -c . $O/r2s8_metrics.csv
echo "=== f64: DMMA vs CUDA cores"; timeout 600 python tools/f64_probe.py 4096 8192 2>&1 | tee $O/r2s8_f64.jsonl
echo "=== layers bench"; timeout 600 python tools/layers_bench.py 2>&1 | tee $O/r2s8_layers_bench.txt | grep -E "conv2d|im2col|check"
