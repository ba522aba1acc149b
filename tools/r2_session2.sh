#!/bin/bash
# Round 2, GPU session 2: first silicon run of the unified kernel (combined stages, dynamic scheduler, PDL), fused prep.
mkdir -p gpurun_out
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 || { echo SMOKE_FAILED; }
echo "=== quick probes"; for v in "LASER_B200_DYNSCHED=1 LASER_B200_PDL=1" "LASER_B200_DYNSCHED=0 LASER_B200_PDL=1" "LASER_B200_DYNSCHED=1 LASER_B200_PDL=0"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>gpurun_out/r2s2_err.log | cut -c1-1400; done
for kc in 128 512; do echo "--- KC=$kc"; LASER_B200_KC=$kc timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>gpurun_out/r2s2_err.log | cut -c1-1400; done
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2s2_pytest.log
echo "=== layouts"; timeout 400 python tools/r2_probe_f16.py 2>&1 | tee gpurun_out/r2s2_layouts.log
echo "=== ncu metrics (single pass)"; timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"gemm_tc_kernel|absmax_mn|split_rows_f16x2|f16x2_rows" -c 12 --csv --log-file gpurun_out/r2s2_dram.csv python tools/r2_ncu_f16_target.py > gpurun_out/r2s2_dram.log 2>&1; grep -c . gpurun_out/r2s2_dram.csv
echo "=== ncu full (rep converted to csv on the box, rep deleted: gpurun_out must stay under 64 MiB)"
NCU_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o /tmp/r2s2_full python tools/r2_ncu_f16_target.py > gpurun_out/r2s2_full.log 2>&1; tail -2 gpurun_out/r2s2_full.log
ncu -i /tmp/r2s2_full.ncu-rep --page raw --csv > gpurun_out/r2s2_full_raw.csv 2>/dev/null
ncu -i /tmp/r2s2_full.ncu-rep --page details --csv > gpurun_out/r2s2_full_details.csv 2>/dev/null
ncu -i /tmp/r2s2_full.ncu-rep --page source --csv > gpurun_out/r2s2_full_source.csv 2>/dev/null
NCU_REPS=1 timeout 300 ncu --set full --clock-control none -k regex:"absmax_mn|split_rows_f16x2|f16x2_rows" -c 3 -o /tmp/r2s2_prep python tools/r2_ncu_f16_target.py > gpurun_out/r2s2_prep.log 2>&1
ncu -i /tmp/r2s2_prep.ncu-rep --page raw --csv > gpurun_out/r2s2_prep_raw.csv 2>/dev/null
gzip -f gpurun_out/r2s2_full_source.csv
du -sh gpurun_out; ls -la gpurun_out | grep r2s2
