#!/bin/bash
# per-kernel launch list (ncu, gpu__time_duration) of the SIL-bearing metric shape and BASELINE configs 3 / 4; then the same un-profiled
mkdir -p gpurun_out
if [ "$1" != "--no-tests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
for w in sil 3 4; do
  STEPS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python scripts/variant_launches.py $w > gpurun_out/variant_$w.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$w.csv')) if len(r)>10 and r[0].isdigit()]
print("== $w", len(rows), "launches")
for r in rows[-(len(rows)//2):]: print(r[4][:70], r[7], r[8], r[-1])
PY
done
python scripts/variant_launches.py
