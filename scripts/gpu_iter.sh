#!/bin/bash
# quick iteration: parity tests + device-timed bench + launch list + phase timers (if the prof build exists)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
tail -3 gpurun_out/bench_quick.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('value %.3e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[7], r[8], r[-1])
PY
if [ -f bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so ]; then
BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so timeout 300 python scripts/phase_prof.py 4096 > gpurun_out/phase.log 2>&1
cat gpurun_out/phase.log
fi
