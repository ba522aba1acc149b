#!/bin/bash
# Round 2, GPU session 5: column factors of the scaled epilogue fetched at tile start (shuffles at store time); host-pointer
# entry panel geometry; raw PCIe rates.
mkdir -p gpurun_out
O=gpurun_out
echo "=== probes"
for v in "X=1" "X=2" "LASER_B200_KC=256"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>$O/r2s5_err.log | tee -a $O/r2s5_probes.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.3f kernel %.3f prep %.3f mre %.2e' % (d['ms'],d['kernel_ms'],d['prep_ms_per_step'],d['error_vs_fp64_S_U(-0.1,0.1)']['f16x3']['mean_relative_error']))"; done
echo "=== pytest (tensor-core files)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zy_f16x3_mode.py tests/test_gpu_fused_epilogue.py tests/test_gpu_prepacked.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r2s5_pytest.log
echo "=== pcie"; timeout 120 python - <<'PY' 2>&1 | tee $O/r2s5_pcie.txt
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both at once: %.1f GB/s per direction (256 MiB pinned)" % (n / a / 1e9, n / b / 1e9, n / c / 1e9))
PY
echo "=== e2e panel geometry"
for v in "X=1" "LASER_B200_PANEL_TAPER=1" "LASER_B200_PANEL_ROWS=512" "LASER_B200_PANEL_ROWS=512 LASER_B200_PANEL_TAPER=1" "LASER_B200_PANEL_ROWS=2048 LASER_B200_PANEL_TAPER=1"; do
  env $v timeout 200 python tools/e2e_probe.py 2>>$O/r2s5_err.log | tee -a $O/r2s5_e2e.jsonl; done
echo "=== ncu metrics (single pass)"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"gemm_tc_kernel" -c 2 --csv --log-file $O/r2s5_metrics.csv python tools/r2_ncu_f16_target.py > $O/r2s5_metrics.log 2>&1; grep -E "time_duration|tensor_cycles|per_second" $O/r2s5_metrics.csv | cut -d, -f5,13,15 | tail -8
