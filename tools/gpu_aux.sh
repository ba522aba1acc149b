#!/bin/bash
mkdir -p gpurun_out
echo "=== fuzz"; timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -6
echo "=== ncu aux"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_simt|gemv_warp|pack_general" -c 6 -o gpurun_out/prof_aux2 python tools/ncu_aux_target.py > gpurun_out/ncu_aux2.log 2>&1; tail -1 gpurun_out/ncu_aux2.log
