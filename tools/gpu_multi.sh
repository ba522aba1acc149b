#!/bin/bash
# usage: gpu_multi.sh N   (under gpurun --gpus N): multi-GPU tests, the row-shard probe and bench.py --gpus N
N=$1
mkdir -p gpurun_out
echo "=== pytest multi-GPU files"; NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_rowshard.py tests/test_c_harness.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_multi_pytest_n$N.log
echo "=== rowshard probe N=$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/rowshard_probe.py > gpurun_out/r2_probe_n$N.log 2>&1; grep -E "world=" gpurun_out/r2_probe_n$N.log; grep -iE "error|signal|exitcode|Traceback" gpurun_out/r2_probe_n$N.log | head -8
echo "=== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n${N}_err.log; head -c 1500 gpurun_out/r2_bench_n$N.json; echo; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n$N.json").read())
    print("N=%d value %.1f ms %.3f | strong %.1f TFLOP/s ms %.3f | parity %s %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["strong_m32768"]["value"], d["strong_m32768"]["ms_per_step"], d["parity"]["ok"], d["strong_m32768"]["parity"]["ok"] if d["strong_m32768"]["parity"] else None))
except Exception as e:
    print("bench line unreadable:", e)
PY
grep -iE "error|signal|exitcode|Traceback|PARITY" gpurun_out/r2_bench_n${N}_err.log | head -8
