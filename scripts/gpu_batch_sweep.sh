#!/bin/bash
mkdir -p gpurun_out
for b in 1024 2048 4096 8192 16384; do
  timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-e2e --no-cpu-baseline > gpurun_out/bench_b$b.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_b$b.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("B=$b value %.3e ms/step %.4f kernel_ms %.4f frac %.3f"%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
else:
    print("B=$b failed"); print(open('gpurun_out/bench_b$b.log').read()[-800:])
PY
done
