"""Print the SASS instructions of an ncu --page source --csv dump that collected the most stall samples, grouped into
contiguous regions, with the dominant stall reason of each."""
import csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_band_source.csv"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = [dict(zip(h, r)) for r in rows[hi + 1:] if len(r) == len(h)]
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(d["# Samples"]) for d in data)
print("total samples", tot, "instructions", len(data))
def dom(d):
    v = sorted(((int(d[c]), c) for c in stalls), reverse=True)[:2]
    return " ".join(f"{c[6:]}:{n}" for n, c in v if n)
if "--regions" in sys.argv:
    # regions of 32 instructions
    R = 32
    for i in range(0, len(data), R):
        blk = data[i:i + R]
        n = sum(int(d["# Samples"]) for d in blk)
        ex = sum(int(d["Instructions Executed"]) for d in blk)
        if n * 200 > tot:
            agg = {c: sum(int(d[c]) for d in blk) for c in stalls}
            v = sorted(((n_, c) for c, n_ in agg.items()), reverse=True)[:3]
            print(f"{i:5d} {100*n/tot:5.1f}% exec {ex:10d}  " + " ".join(f"{c[6:]}:{n_}" for n_, c in v) + "   | " + blk[0]["Source"].strip()[:50])
else:
    order = sorted(range(len(data)), key=lambda i: -int(data[i]["# Samples"]))[:top]
    for i in sorted(order):
        d = data[i]
        print(f"{i:5d} {100*int(d['# Samples'])/tot:5.2f}% ex {d['Instructions Executed']:>9s}  {d['Source'].strip()[:70]:70s} {dom(d)}")
