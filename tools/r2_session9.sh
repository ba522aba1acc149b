#!/bin/bash
# Round 2, GPU session 9: C through smem-staged TMA stores (UTMASTG) against plain stores; beta != 0; GPU suite.
mkdir -p gpurun_out
O=gpurun_out
echo "=== probes"
for v in "X=1" "LASER_B200_C_TMA=0" "X=2" "LASER_B200_C_TMA=0"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>$O/r2s9_err.log | tee -a $O/r2s9_probes.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.3f kernel %.3f prep %.3f mre %.2e' % (d['ms'],d['kernel_ms'],d['prep_ms_per_step'],d['error_vs_fp64_S_U(-0.1,0.1)']['f16x3']['mean_relative_error']))"; done
echo "=== beta / modes"; for v in "X=1" "LASER_B200_C_TMA=0"; do env $v timeout 300 python tools/panel_probe.py 2>&1 | grep -E "full K=8192" ; done
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2s9_pytest.log
echo "=== ncu metrics (single pass)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum
for v in "X=1" "LASER_B200_C_TMA=0"; do env $v NCU_REPS=2 timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel" -c 2 --csv --log-file $O/r2s9_m.csv python tools/r2_ncu_f16_target.py > /dev/null 2>&1; echo "--- $v"; grep gemm_tc $O/r2s9_m.csv | cut -d, -f13,15 | tr '\n' ' '; echo; done
