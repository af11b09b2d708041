#!/bin/bash
# ncu --set full of the planner and the stamp/confidence kernel (one launch each, after warm-up)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"plan_kernel|assort_confidence_kernel" -s 6 -c 2 -f -o gpurun_out/prof_small \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_small_run.log 2>&1
ls -la gpurun_out/prof_small.ncu-rep
