#!/bin/bash
# usage: gpu_multi.sh N   (under gpurun --gpus N): multi-GPU tests, the row-shard probe (B prepared in 2 / 4 / 1 panels, raw
# broadcast) and bench.py --gpus N
N=$1
mkdir -p gpurun_out
echo "=== pytest multi-GPU files"; NCCL_DEBUG=WARN timeout 1200 python -m pytest tests/test_gpu_rowshard.py tests/test_c_harness.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_multi_pytest_n$N.log
for P in 2 4 1 0; do
echo "=== rowshard probe N=$N LASER_B200_ROWSHARD_PANELS=$P"; LASER_B200_ROWSHARD_PANELS=$P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$P tools/rowshard_probe.py > gpurun_out/r2_probe_n${N}_p$P.log 2>&1; grep -E "world=" gpurun_out/r2_probe_n${N}_p$P.log; grep -iE "error|signal|exitcode|Traceback" gpurun_out/r2_probe_n${N}_p$P.log | head -4
done
echo "=== bench N=$N"; bash tools/gpu_bench_n.sh $N
