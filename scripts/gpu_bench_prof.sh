#!/bin/bash
# bench + ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -25 gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi -s 3 -c 2 -f -o gpurun_out/prof_viterbi \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/
