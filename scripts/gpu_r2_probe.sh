#!/bin/bash
# round 2, first probe: launch-overlap and device-side launch feasibility + today's baseline of the round-1 build
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for w in -1 75; do for t in 0 1; do
  timeout 60 scripts/micro/pdl_chain 20 130 8 $w $t
done; done
timeout 60 scripts/micro/pdl_chain 20 130 0 -1 0
timeout 60 scripts/micro/cdp_tail
} > gpurun_out/r2_probe.log 2>&1
cat gpurun_out/r2_probe.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
tail -2 gpurun_out/bench_quick.log | cut -c1-1500
