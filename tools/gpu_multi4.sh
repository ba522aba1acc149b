#!/bin/bash
# under gpurun --gpus 4: row-shard probe with B prepared in 1 / 2 panels, then bench.py --gpus 4
mkdir -p gpurun_out
for P in 1 2; do
echo "=== rowshard probe N=4 LASER_B200_ROWSHARD_PANELS=$P"; LASER_B200_ROWSHARD_PANELS=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2957$P tools/rowshard_probe.py > gpurun_out/r2_probe_n4_p$P.log 2>&1; grep -E "world=" gpurun_out/r2_probe_n4_p$P.log; grep -iE "error|Traceback" gpurun_out/r2_probe_n4_p$P.log | head -3
done
bash tools/gpu_bench_n.sh 4
