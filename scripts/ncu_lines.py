"""Aggregate the stall samples of an ncu source-page dump (SASS view, --csv) by CUDA source line.
The line of every SASS instruction comes from `nvdisasm -g -c` of the cubin inside libbfa_b200.so (same build!).
usage: ncu_lines.py <source.csv> <mangled kernel substring> [min_pct]"""
import csv, re, subprocess, sys, tempfile, os, collections
src_csv = sys.argv[1]; pat = sys.argv[2]; min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
lib = "bournemouth-forced-aligner_b200/lib/libbfa_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
txt = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if pat in t: txt = t; break
assert txt, "kernel not found"
# locate the function
i0 = txt.index(".text." + [m for m in re.findall(r"\.text\.(\S+)", txt) if pat in m][0])
lines = txt[i0:].split("\n")
sass_line = []   # source line of each instruction in order
cur = None
for l in lines[1:]:
    if l.startswith("\t.section") or l.startswith(".section"): 
        if sass_line: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): sass_line.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = [dict(zip(h, r)) for r in rows[hi + 1:] if len(r) == len(h)]
print("sass in cubin", len(sass_line), "in report", len(data))
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: collections.Counter())
for d, sl in zip(data, sass_line):
    a = agg[sl]; a["n"] += int(d["# Samples"]); a["ex"] += int(d["Instructions Executed"])
    for c in stalls: a[c] += int(d[c])
tot = sum(a["n"] for a in agg.values())
srcs = {}
def src(fl):
    if fl is None: return ""
    f, n = fl
    if f not in srcs:
        p = os.path.join("bournemouth-forced-aligner_b200/csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return srcs[f][n - 1].strip()[:80] if n - 1 < len(srcs[f]) else ""
for fl in sorted(agg, key=lambda x: (x is None, x)):
    a = agg[fl]
    if a["n"] * 100 >= min_pct * tot:
        v = sorted(((a[c], c) for c in stalls), reverse=True)[:3]
        print(f"{str(fl[1]) if fl else '-':>5} {100*a['n']/tot:5.1f}% ex {a['ex']:>9} {src(fl):80s} " + " ".join(f"{c[6:]}:{n}" for n, c in v if n))
