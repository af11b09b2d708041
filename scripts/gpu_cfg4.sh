STEPS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_4.csv python scripts/variant_launches.py 4 > gpurun_out/variant_4.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_4.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-(len(rows)//2):]: print(r[4][:70], r[7], r[8], r[-1])
PY
