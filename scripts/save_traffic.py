"""Refresh profiles/traffic_latest.json from gpurun_out/prof_direct.ncu-rep (ncu --set full of the direct kernel taken at THIS build:
the file carries the hash of the kernel sources, bench.py quotes roofline.traffic only when it matches the build it runs)."""
import csv, hashlib, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "latest"
h = hashlib.sha256()
for f in sorted((ROOT / "bournemouth-forced-aligner_b200" / "csrc").glob("*.cu*")):
    h.update(f.name.encode()); h.update(f.read_bytes())
rep = ROOT / "gpurun_out" / "prof_direct.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
m = {n: (u, v) for n, u, v in zip(rows[0], rows[1], rows[2])}
def num(k):
    u, v = m[k]; x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
tr = {"kernel": m["Kernel Name"][1], "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
      "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
      "gpu_time_duration_us_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")),
      "inst_executed": float(m["smsp__inst_executed.sum"][1].replace(",", "")), "csrc_sha16": h.hexdigest()[:16],
      "source": f"profiles/{tag}_direct_ncu_details.csv (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3, launch 7 of the kernel)"}
(ROOT / "profiles" / "traffic_latest.json").write_text(json.dumps(tr, indent=1) + "\n")
(ROOT / "profiles" / f"{tag}_direct_traffic.json").write_text(json.dumps(tr, indent=1) + "\n")
(ROOT / "profiles" / f"{tag}_direct_ncu_details.csv").write_text(
    subprocess.run(["ncu", "-i", str(rep), "--page", "details", "--csv"], capture_output=True, text=True).stdout)
print(json.dumps(tr))
