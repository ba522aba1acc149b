#!/bin/bash
# Round 2, GPU session 14: A/B (LASER_B200_LIB) of the build with bias / activation out of line (tensor-core kernel 374 -> 176 KB,
# exact SIMT kernel 133 -> 76 KB of code) against the previous one.
mkdir -p gpurun_out
PREV=laser_b200/lib/prev/liblaser_b200_prev.so
for i in 1 2 3; do
for v in "X=1" "LASER_B200_LIB=$PREV"; do
  echo "--- $v"; env $v timeout 200 python tools/two_piece_probe.py f16x3 8192 10 2>>gpurun_out/r2s14_err.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.3f kernel %.3f prep %.3f' % (d['ms'],d['kernel_ms'],d['prep_ms_per_step']))"; done; done
echo "=== exact SIMT 4096^3"
for v in "X=1" "LASER_B200_LIB=$PREV"; do env $v timeout 200 python - <<'PY'
import os, torch, laser_b200 as L
L.init(); n = 4096
a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
f = lambda: L.gemm_strided(n, n, n, 1.0, a, n, 1, b, n, 1, 0.0, c, n, 1, path=L.PATH_SIMT)
for _ in range(2): f()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
print(os.environ.get("LASER_B200_LIB", "new"), "simt 4096^3 %.3f ms %.1f TFLOP/s" % (ms, 2 * n**3 / ms / 1e9))
PY
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for v in "X=1" "LASER_B200_LIB=$PREV"; do env $v NCU_REPS=2 timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_tc_kernel" -c 2 --csv --log-file gpurun_out/r2s14_m.csv python tools/r2_ncu_f16_target.py > /dev/null 2>&1; echo "--- $v"; grep gemm_tc gpurun_out/r2s14_m.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' '; echo; done
echo "=== pytest (fused epilogue + parity)"; timeout 900 python -m pytest tests/test_gpu_fused_epilogue.py tests/test_gpu_parity.py tests/test_gpu_zlayers.py -m gpu -x -q 2>&1 | tail -3
