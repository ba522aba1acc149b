#!/bin/bash
# the last GPU call of the round: smoke, the whole GPU suite and one bench line on the final build
mkdir -p gpurun_out
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_final_bench_err.log | tee gpurun_out/r2_final_bench_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.1f ms %.3f | kernel %.3f frac %.3f (sustained %.3f) prep %.3f | e2e %.1f | strong %.1f | parity %s' % (d['value'],d['ms_per_step'],r['kernel_ms'],r['frac'],r['frac_of_sustained_peak'],r['prep_ms_per_step'],d['e2e']['value'],d['strong_m32768']['value'],d['parity']['ok'])); print({k:round(v['tflops'],1) for k,v in d['modes'].items()})"
tail -2 gpurun_out/r2_final_bench_err.log
