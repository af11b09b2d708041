"""Copy the evidence of the last gpu_round.sh run from gpurun_out/ into profiles/ under a tag, and refresh
profiles/traffic_latest.json (DRAM bytes per launch of the dominant kernel from the ncu --set full capture)."""
import csv, json, shutil, subprocess, sys
from pathlib import Path
tag = sys.argv[1]
out = Path("profiles"); g = Path("gpurun_out")
line = [l for l in (g / "bench.log").read_text().splitlines() if l.startswith("{")][-1]
(out / f"{tag}_bench.json").write_text(line + "\n")
ref = [l for l in (g / "bench_ref.log").read_text().splitlines() if l.startswith("{")]
if ref: (out / f"{tag}_bench_reference_arm.json").write_text(ref[-1] + "\n")
shutil.copy(g / "launches.csv", out / f"{tag}_launches.csv")
det = subprocess.run(["ncu", "-i", str(g / "prof_band.ncu-rep"), "--page", "details", "--csv"], capture_output=True, text=True).stdout
(out / f"{tag}_band_ncu_details.csv").write_text(det)
raw = subprocess.run(["ncu", "-i", str(g / "prof_band.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (uu, vv) for n, uu, vv in zip(h, u, v)}
def mb(k):
    uu, vv = m[k]; x = float(vv.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[uu]
tr = {"kernel": m["Kernel Name"][1] if "Kernel Name" in m else "", "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum"),
      "dram_bytes_per_launch": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
      "gpu_time_duration_us_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")),
      "inst_executed": float(m["smsp__inst_executed.sum"][1].replace(",", "")), "source": f"profiles/{tag}_band_ncu_details.csv (ncu --set full --clock-control none)"}
(out / "traffic_latest.json").write_text(json.dumps(tr, indent=1) + "\n")
(out / f"{tag}_band_traffic.json").write_text(json.dumps(tr, indent=1) + "\n")
print(json.dumps(tr))
