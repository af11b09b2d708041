#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_band -s 6 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
