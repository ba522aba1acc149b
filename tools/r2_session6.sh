#!/bin/bash
# Round 2, GPU session 6: split-K of the last partial wave (4096^3: 3.5 waves instead of 4), full GPU suite, bench, ncu of the
# shuffled-epilogue kernel, per-layout DRAM / L2 / tensor metrics.
mkdir -p gpurun_out
O=gpurun_out
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | grep -E "f16x3|M=32768" | tee $O/r2s6_layouts.log
echo "=== 4096 split-K of the tail off"; LASER_B200_SPLITK=0 timeout 200 python tools/r2_probe_f16.py 2>&1 | grep -E "n=4096 f16x3" | tee $O/r2s6_layouts_nosplit.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2s6_pytest.log
echo "=== bench reference arm"; LASER_B200_REF_BUDGET_S=12 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/r2s6_bench_ref_err.log | tee $O/r2s6_bench_ref.json | cut -c1-300
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>$O/r2s6_bench_err.log | tee $O/r2s6_bench_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.1f ms %.3f | kernel %.3f frac %.3f prep %.3f | e2e %.1f (%.2f ms) | strong %.1f | parity %s' % (d['value'],d['ms_per_step'],r['kernel_ms'],r['frac'],r['prep_ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['strong_m32768']['value'],d['parity']['ok'])); print({k:round(v['tflops'],1) for k,v in d['modes'].items()}); print(d['clocks'])"
tail -3 $O/r2s6_bench_err.log
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2s6_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/r2s6_bench_under_ncu.log 2>&1; grep -c . $O/r2s6_launches_bench.csv
echo "=== ncu metrics (single pass)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows" -c 8 --csv --log-file $O/r2s6_metrics.csv python tools/r2_ncu_f16_target.py > $O/r2s6_metrics.log 2>&1; grep -c . $O/r2s6_metrics.csv
timeout 300 ncu --metrics $M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_utcmma.sum --clock-control none -k regex:"gemm_tc_kernel" -s 4 -c 4 --csv --log-file $O/r2s6_layout_metrics.csv python tools/r2_ncu_layout_target.py > $O/r2s6_layout_metrics.log 2>&1; grep -c . $O/r2s6_layout_metrics.csv
echo "=== ncu full"
NCU_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o /tmp/r2s6_full python tools/r2_ncu_f16_target.py > $O/r2s6_full.log 2>&1; tail -2 $O/r2s6_full.log
ncu -i /tmp/r2s6_full.ncu-rep --page raw --csv > $O/r2s6_full_raw.csv 2>/dev/null
ncu -i /tmp/r2s6_full.ncu-rep --page details --csv > $O/r2s6_full_details.csv 2>/dev/null
ncu -i /tmp/r2s6_full.ncu-rep --page source --csv > $O/r2s6_full_source.csv 2>/dev/null; gzip -f $O/r2s6_full_source.csv
du -sh $O
