#!/bin/bash
# Round 2, GPU session 4 (results of sessions 1-3 were lost with their container): the full evidence pass of the current
# build -- smoke, GPU suite, probes, both bench arms, ncu launch list of the bench command, single-pass DRAM metrics,
# --set full captures (converted to csv on the box: gpurun_out must stay under 64 MiB).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/r2s4_smi.txt
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== probes"
for v in "X=1" "LASER_B200_DYNSCHED=0" "LASER_B200_PDL=0" "LASER_B200_KC=256" "LASER_B200_RASTER=4" "LASER_B200_RASTER=16"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>$O/r2s4_err.log | tee -a $O/r2s4_probes.jsonl | cut -c1-1200; done
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | tee $O/r2s4_layouts.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r2s4_pytest.log
echo "=== bench reference arm"; LASER_B200_REF_BUDGET_S=12 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/r2s4_bench_ref_err.log | tee $O/r2s4_bench_ref.json | cut -c1-600
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>$O/r2s4_bench_err.log | tee $O/r2s4_bench_n1.json | cut -c1-4000
tail -3 $O/r2s4_bench_err.log
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2s4_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/r2s4_bench_under_ncu.log 2>&1; grep -c . $O/r2s4_launches_bench.csv
echo "=== ncu metrics (single pass)"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows" -c 9 --csv --log-file $O/r2s4_metrics.csv python tools/r2_ncu_f16_target.py > $O/r2s4_metrics.log 2>&1; grep -c . $O/r2s4_metrics.csv
echo "=== ncu full"
NCU_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o /tmp/r2s4_full python tools/r2_ncu_f16_target.py > $O/r2s4_full.log 2>&1; tail -2 $O/r2s4_full.log
ncu -i /tmp/r2s4_full.ncu-rep --page raw --csv > $O/r2s4_full_raw.csv 2>/dev/null
ncu -i /tmp/r2s4_full.ncu-rep --page details --csv > $O/r2s4_full_details.csv 2>/dev/null
ncu -i /tmp/r2s4_full.ncu-rep --page source --csv > $O/r2s4_full_source.csv 2>/dev/null; gzip -f $O/r2s4_full_source.csv
NCU_REPS=1 timeout 300 ncu --set full --clock-control none -k regex:"absmax_mn|split_rows_f16x2|f16x2_rows" -c 2 -o /tmp/r2s4_prep python tools/r2_ncu_f16_target.py > $O/r2s4_prep.log 2>&1
ncu -i /tmp/r2s4_prep.ncu-rep --page raw --csv > $O/r2s4_prep_raw.csv 2>/dev/null
ncu -i /tmp/r2s4_prep.ncu-rep --page details --csv > $O/r2s4_prep_details.csv 2>/dev/null
du -sh $O; ls -la $O | grep r2s4
