#!/bin/bash
# iteration run: parity tests, bench, ncu of the hot kernel + assort kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_band -s 6 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assort -s 3 -c 1 -f -o gpurun_out/prof_assort \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run2.log 2>&1
ls -la gpurun_out/*.ncu-rep
