#!/bin/bash
# quick iteration: parity tests + bench (device-timed only) + ncu full capture of the band kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
tail -3 gpurun_out/bench_quick.log | cut -c1-1500
if [ "$1" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_band -s 6 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
