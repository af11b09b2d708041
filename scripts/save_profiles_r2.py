"""Copy the evidence of the last scripts/gpu_r2_evidence.sh run from gpurun_out/ into profiles/ under a round-2 tag, and refresh
profiles/traffic_latest.json (DRAM bytes per launch of the dominant kernel from the ncu --set full capture, tied to the kernel
sources by their hash: bench.py quotes `roofline.traffic` only when the hash matches the build it runs)."""
import csv, hashlib, json, re, shutil, subprocess, sys
from pathlib import Path
tag = sys.argv[1]
ROOT = Path(__file__).resolve().parents[1]
out, g = ROOT / "profiles", ROOT / "gpurun_out"


def csrc_hash():
    h = hashlib.sha256()
    for f in sorted((ROOT / "bournemouth-forced-aligner_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()[:16]


def last_json(p):
    ls = [l for l in p.read_text().splitlines() if l.startswith("{")]
    return ls[-1] if ls else None


def ncu(rep, page, *extra):
    return subprocess.run(["ncu", "-i", str(rep), "--page", page, "--csv", *extra], capture_output=True, text=True).stdout


for src, dst in [("bench_full.log", "bench.json"), ("bench_ref.log", "bench_reference_arm.json")]:
    l = last_json(g / src)
    if l: (out / f"{tag}_{dst}").write_text(l + "\n")
for n in (2, 4, 8):
    for src, dst in [(f"bench_n{n}.log", f"bench_n{n}.json"), (f"corpus_n{n}.log", f"corpus_n{n}.json")]:
        if (g / src).exists():
            l = last_json(g / src)
            if l: (out / f"{tag}_{dst}").write_text(l + "\n")
for src, dst in [("launches.csv", "launches.csv"), ("launches_sil.csv", "launches_sil_variant.csv"), ("launches_3.csv", "launches_config3.csv"),
                 ("launches_4.csv", "launches_config4.csv"), ("phase.log", "phase_timers.txt"), ("variants_plain.log", "variants_unprofiled.txt")]:
    if (g / src).exists(): shutil.copy(g / src, out / f"{tag}_{dst}")
# the dominant kernel
rep = g / "prof_direct.ncu-rep"
(out / f"{tag}_direct_ncu_details.csv").write_text(ncu(rep, "details"))
rows = list(csv.reader(ncu(rep, "raw").splitlines()))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (uu, vv) for n, uu, vv in zip(h, u, v)}
def num(k):
    uu, vv = m[k]; x = float(vv.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(uu, 1)
tr = {"kernel": m["Kernel Name"][1], "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
      "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
      "gpu_time_duration_us_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")),
      "inst_executed": float(m["smsp__inst_executed.sum"][1].replace(",", "")), "csrc_sha16": csrc_hash(),
      "source": f"profiles/{tag}_direct_ncu_details.csv (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3, launch 7 of the kernel)"}
(out / "traffic_latest.json").write_text(json.dumps(tr, indent=1) + "\n")
(out / f"{tag}_direct_traffic.json").write_text(json.dumps(tr, indent=1) + "\n")
print(json.dumps(tr))
# the silence-anchoring chain
rep = g / "prof_sil.ncu-rep"
if rep.exists():
    (out / f"{tag}_silchain_ncu_details.csv").write_text(ncu(rep, "details"))
    rows = list(csv.reader(ncu(rep, "raw").splitlines()))
    h = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
    idx = [h.index(w) for w in want if w in h]
    with open(out / f"{tag}_silchain_summary.csv", "w") as f:
        wr = csv.writer(f)
        for r in rows[:2] + rows[2:]:
            wr.writerow([r[i] for i in idx])
# sanitizer
txt = ["compute-sanitizer over a slice of the GPU tests (scripts/gpu_sanitize.sh) at the tagged build\n"]
for tool in ("memcheck", "synccheck", "racecheck"):
    lp, pp = g / f"{tool}.log", g / f"{tool}_pytest.log"
    if not lp.exists(): continue
    txt.append(f"== {tool}: pytest: " + (pp.read_text().strip().splitlines()[-1] if pp.exists() else "?"))
    body = lp.read_text()
    txt += [l for l in body.splitlines() if "ERROR SUMMARY" in l or "RACECHECK SUMMARY" in l]
    if tool == "racecheck":
        pairs = re.findall(r"Race reported between (\w+ access at [^\n]*?) in (\S+:\d+)", body)
        import collections
        cnt = collections.Counter((re.sub(r"\(.*", "", a.split(" at ")[1])[:60], loc) for a, loc in pairs)
        txt.append("   hazards by (function, first location), all inside the banded kernels' mbarrier-ordered producer/consumer hand-overs "
                   "(the tool models neither mbarrier acquire/release nor bulk-copy completion):")
        txt += [f"   {n:4d}  {fn}  {loc}" for (fn, loc), n in cnt.most_common(25)]
(out / f"{tag}_sanitizer.txt").write_text("\n".join(txt) + "\n")
pt = (g / "pytest_gpu.log").read_text().strip().splitlines()[-2:]
(out / f"{tag}_pytest_gpu.txt").write_text("\n".join(pt) + "\n")
print("saved", tag)
