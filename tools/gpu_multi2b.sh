#!/bin/bash
# under gpurun --gpus 2: multi-GPU tests + probe of the double-buffered prepared-panel path
mkdir -p gpurun_out
echo "=== pytest multi-GPU files"; NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_rowshard.py -m gpu -x -q 2>&1 | tail -4
for P in 1 2; do
echo "=== rowshard probe N=2 LASER_B200_ROWSHARD_PANELS=$P"; LASER_B200_ROWSHARD_PANELS=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2958$P tools/rowshard_probe.py > gpurun_out/r2_probe2_n2_p$P.log 2>&1; grep -E "world=" gpurun_out/r2_probe2_n2_p$P.log; grep -iE "error|Traceback" gpurun_out/r2_probe2_n2_p$P.log | head -3
done
