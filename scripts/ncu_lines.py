"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
stall samples and executed instructions per line, top N."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
# The export lists source lines first (with aggregated metrics) when print-source is cuda,sass; take rows whose first col is int and Address empty
col = {n: i for i, n in enumerate(h)}
samples = collections.Counter(); insts = collections.Counter(); text = {}
cur_file = ""
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) < len(h): 
        if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
        continue
    try: ln = int(r[0])
    except: continue
    addr = r[2]
    if addr not in ("", "-"): continue   # SASS rows
    s = int(r[col["# Samples"]] or 0); n = int(r[col["Instructions Executed"]] or 0)
    key = (cur_file, ln)
    samples[key] += s; insts[key] += n; text[key] = r[1].strip()[:110]
    tot_s += s; tot_i += n
print("total samples", tot_s, "total inst", tot_i)
print("--- by samples")
for k, s in samples.most_common(topn):
    print(f"{k[0]}:{k[1]:4d} samp {s:6d} ({100*s/max(tot_s,1):4.1f}%) inst {insts[k]:9d} ({100*insts[k]/max(tot_i,1):4.1f}%) | {text[k]}")
if len(sys.argv) > 3:
    # region summary: "name:lo-hi,name:lo-hi"
    print("--- regions (viterbi file only)")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        s = sum(v for k, v in samples.items() if k[0].startswith("viterbi") and lo <= k[1] <= hi)
        n = sum(v for k, v in insts.items() if k[0].startswith("viterbi") and lo <= k[1] <= hi)
        print(f"{name:12s} samples {100*s/tot_s:5.1f}%  inst {100*n/tot_i:5.1f}%")
    s = sum(v for k, v in samples.items() if not k[0].startswith("viterbi")); n = sum(v for k, v in insts.items() if not k[0].startswith("viterbi"))
    print(f"{'other files':12s} samples {100*s/tot_s:5.1f}%  inst {100*n/tot_i:5.1f}%")
