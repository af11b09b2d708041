#!/bin/bash
mkdir -p gpurun_out


for f in "" "--unfused-conf"; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline $f > gpurun_out/bench_ab.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_ab.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("$f value %.3e step %.4f ms kernel %.4f ms frac %.3f"%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
else: print(open('gpurun_out/bench_ab.log').read()[-1500:])
PY
done
