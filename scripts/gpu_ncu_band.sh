#!/bin/bash
# one ncu --set full capture of the metric-shape banded kernel (<3,66> is launch 0 of every 3 band launches per step)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:band3_kernel<.int.66" -s 3 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
