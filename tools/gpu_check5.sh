#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "=== rowshard test"; timeout 900 python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x 2>&1 | tail -5
echo "=== bench n=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench2_err.log | tee gpurun_out/bench_n2.json | cut -c1-1500; tail -3 gpurun_out/bench2_err.log
echo "=== bench n=1"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-300
