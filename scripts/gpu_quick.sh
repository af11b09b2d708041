#!/bin/bash
# quick iteration: parity tests + bench (device-timed only) + per-kernel launch list (+ optional ncu full capture)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
tail -3 gpurun_out/bench_quick.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-7:]: print(r[4][:60], r[7], r[8], r[-1])
PY
if [ "$1" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_band3_kernel.3 -s 6 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
