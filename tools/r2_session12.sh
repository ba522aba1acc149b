#!/bin/bash
# Round 2, GPU session 12: cp.async-staged few-rows kernel (im2col convolution), layers bench, layer + parity tests, sanitizer.
mkdir -p gpurun_out
O=gpurun_out
echo "=== layers bench"; timeout 600 python tools/layers_bench.py 2>&1 | tee $O/r2s12_layers_bench.txt | grep -E "conv2d|im2col|check"
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_zlayers.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "=== sanitizer"; timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_target.py 2>&1 | grep -E "ERROR SUMMARY|few rows|done" | tail -4
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_skinny|im2col" -c 4 --csv --log-file $O/r2s12_m.csv python tools/ncu_r2_aux_target.py > /dev/null 2>&1; grep -E "skinny|im2col" $O/r2s12_m.csv | awk -F'","' '{print substr($5,1,40), $13, $15}'
