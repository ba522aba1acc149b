#!/bin/bash
# Round 2, GPU session 10: the evidence pass of the final build -- smoke, GPU suite, compute-sanitizer, probes, both bench
# arms, ncu launch list of the bench command, single-pass DRAM / tensor metrics, --set full captures (converted to csv on the
# box: gpurun_out must stay under 64 MiB), layers and fp64 timings.
mkdir -p gpurun_out
O=gpurun_out
P=r2s10
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/${P}_smi.txt
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/${P}_pytest.log
echo "=== compute-sanitizer"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_target.py > $O/${P}_sanitizer_memcheck.log 2>&1; grep -E "ERROR SUMMARY|done|few rows|dmma|tail split|ring prep" $O/${P}_sanitizer_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_target.py > $O/${P}_sanitizer_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|done" $O/${P}_sanitizer_racecheck.log | tail -4
echo "=== probes"
for v in "X=1" "LASER_B200_KC=256"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>$O/${P}_err.log | tee -a $O/${P}_probes.jsonl | cut -c1-900; done
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | tee $O/${P}_layouts.log
echo "=== large / skewed shapes"; timeout 600 python tools/large_shapes.py 2>&1 | tee $O/${P}_large_shapes.txt | tail -14
echo "=== f64"; timeout 600 python tools/f64_probe.py 2048 4096 8192 2>&1 | tee $O/${P}_f64.jsonl
echo "=== layers bench"; timeout 600 python tools/layers_bench.py 2>&1 | tee $O/${P}_layers_bench.txt | tail -12
echo "=== bench reference arm"; LASER_B200_REF_BUDGET_S=12 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/${P}_bench_ref_err.log | tee $O/${P}_bench_ref.json | cut -c1-300
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>$O/${P}_bench_err.log | tee $O/${P}_bench_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.1f ms %.3f | kernel %.3f frac %.3f prep %.3f | e2e %.1f (%.2f ms) | strong %.1f | parity %s' % (d['value'],d['ms_per_step'],r['kernel_ms'],r['frac'],r['prep_ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['strong_m32768']['value'],d['parity']['ok'])); print({k:round(v['tflops'],1) for k,v in d['modes'].items()}); print(d['clocks'])"
tail -3 $O/${P}_bench_err.log
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/${P}_bench_under_ncu.log 2>&1; grep -c . $O/${P}_launches_bench.csv
echo "=== ncu metrics (single pass)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows" -c 8 --csv --log-file $O/${P}_metrics.csv python tools/r2_ncu_f16_target.py > $O/${P}_metrics.log 2>&1; grep -c . $O/${P}_metrics.csv
echo "=== ncu full"
NCU_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o /tmp/${P}_full python tools/r2_ncu_f16_target.py > $O/${P}_full.log 2>&1; tail -2 $O/${P}_full.log
ncu -i /tmp/${P}_full.ncu-rep --page raw --csv > $O/${P}_full_raw.csv 2>/dev/null
ncu -i /tmp/${P}_full.ncu-rep --page details --csv > $O/${P}_full_details.csv 2>/dev/null
ncu -i /tmp/${P}_full.ncu-rep --page source --csv > $O/${P}_full_source.csv 2>/dev/null; gzip -f $O/${P}_full_source.csv
NCU_REPS=1 timeout 300 ncu --set full --clock-control none -k regex:"absmax_mn|split_rows_f16x2|f16x2_rows" -c 3 -o /tmp/${P}_prep python tools/r2_ncu_f16_target.py > $O/${P}_prep.log 2>&1
ncu -i /tmp/${P}_prep.ncu-rep --page raw --csv > $O/${P}_prep_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"gemm_dmma|gemm_skinny|splitk_tail|im2col|transpose_batched" -c 6 -o /tmp/${P}_aux python tools/ncu_r2_aux_target.py > $O/${P}_aux.log 2>&1
ncu -i /tmp/${P}_aux.ncu-rep --page raw --csv > $O/${P}_aux_raw.csv 2>/dev/null
du -sh $O; ls $O | grep ${P} | wc -l
