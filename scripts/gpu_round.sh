#!/bin/bash
# full evidence run: parity tests, full bench line (e2e + cpu baseline), reference arm, ncu launch list, ncu --set full of the band kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-2500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_ref.log | cut -c1-800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[7], r[8], r[-1])
PY
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:band3_kernel<.int.66" -s 3 -c 1 -f -o gpurun_out/prof_band \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
