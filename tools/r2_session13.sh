#!/bin/bash
# Round 2, GPU session 13: why is the few-rows kernel slow -- full capture with source page
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny" -c 1 -o /tmp/r2s13 python tools/ncu_r2_aux_target.py > gpurun_out/r2s13.log 2>&1; tail -2 gpurun_out/r2s13.log
ncu -i /tmp/r2s13.ncu-rep --page raw --csv > gpurun_out/r2s13_raw.csv 2>/dev/null
ncu -i /tmp/r2s13.ncu-rep --page source --csv > gpurun_out/r2s13_source.csv 2>/dev/null; gzip -f gpurun_out/r2s13_source.csv
