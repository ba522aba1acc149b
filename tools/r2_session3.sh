#!/bin/bash
# Round 2, GPU session 3: packed-conversion prep kernels, full GPU suite, bench.py (both arms), prep kernel timing.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2s3_pytest.log
echo "=== probe"; timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>gpurun_out/r2s3_err.log | tee gpurun_out/r2s3_probe.json | cut -c1-1400
echo "=== ncu metrics (single pass)"; timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows" -c 8 --csv --log-file gpurun_out/r2s3_metrics.csv python tools/r2_ncu_f16_target.py > gpurun_out/r2s3_metrics.log 2>&1; grep -c . gpurun_out/r2s3_metrics.csv
echo "=== bench reference arm"; LASER_B200_REF_BUDGET_S=12 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r2s3_bench_ref_err.log | tee gpurun_out/r2s3_bench_ref.json | cut -c1-900
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2s3_bench_err.log | tee gpurun_out/r2s3_bench_n1.json | cut -c1-3000
tail -5 gpurun_out/r2s3_bench_err.log
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | grep -E "f16x3|M=32768" | tee gpurun_out/r2s3_layouts.log
du -sh gpurun_out
