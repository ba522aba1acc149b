#!/bin/bash
# usage: gpu_multi_panels.sh N   (under gpurun --gpus N): the row-shard probe with B broadcast in 1 / 2 / 4 / 8 pieces
N=$1
mkdir -p gpurun_out
for P in 1 2 4 8; do
  echo "=== LASER_B200_ROWSHARD_PANELS=$P"
  LASER_B200_ROWSHARD_PANELS=$P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$P tools/rowshard_probe.py > gpurun_out/r2_panels_n${N}_p$P.log 2>&1; grep -E "world=" gpurun_out/r2_panels_n${N}_p$P.log; grep -iE "error|Traceback" gpurun_out/r2_panels_n${N}_p$P.log | head -3
done
echo "=== NCCL parity with pieces"; LASER_B200_ROWSHARD_PANELS=4 NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_rowshard.py -m gpu -x -q 2>&1 | tail -3
