#!/bin/bash
# usage: gpu_bench_n.sh N  (under gpurun --gpus N): bench.py --gpus N through torchrun, as the driver launches it
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n${N}_err.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read())
print("N=%d value %.1f ms %.3f | strong %.1f TFLOP/s ms %.3f | e2e %.1f | parity %s %s | kernel %.3f prep %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["strong_m32768"]["value"], d["strong_m32768"]["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["strong_m32768"]["parity"]["ok"] if d["strong_m32768"]["parity"] else None, d["roofline"]["kernel_ms"], d["roofline"]["prep_ms_per_step"]))
PY
grep -iE "error|Traceback|PARITY" gpurun_out/r2_bench_n${N}_err.log | head -5
