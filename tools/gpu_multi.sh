#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
echo "=== rowshard probe N=$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/rowshard_probe.py > gpurun_out/probe_n$N.log 2>&1; grep -E "world=" gpurun_out/probe_n$N.log; grep -iE "error|signal|exitcode|Traceback" gpurun_out/probe_n$N.log | head -8
echo "=== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench${N}_err.log; head -c 600 gpurun_out/bench_n$N.json; grep -iE "error|signal|exitcode|Traceback" gpurun_out/bench${N}_err.log | head -8
echo "=== bench N=$N, A prepared during the broadcast (LASER_B200_ROWSHARD_OVERLAP=1)"; LASER_B200_ROWSHARD_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_overlap_n$N.json 2> gpurun_out/bench_overlap${N}_err.log; head -c 300 gpurun_out/bench_overlap_n$N.json; grep -iE "error|signal|exitcode|Traceback" gpurun_out/bench_overlap${N}_err.log | head -8
echo "=== bench N=$N, B broadcast in 2 K-panels (LASER_B200_ROWSHARD_PANELS=2)"; LASER_B200_ROWSHARD_PANELS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29567 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_panels2_n$N.json 2> gpurun_out/bench_panels2_${N}_err.log; head -c 300 gpurun_out/bench_panels2_n$N.json; grep -iE "error|signal|exitcode|Traceback" gpurun_out/bench_panels2_${N}_err.log | head -8
free -g | head -2; df -h /dev/shm | tail -1
