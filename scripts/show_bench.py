"""Development: print the interesting fields of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = None
for ln in txt.strip().splitlines():
    try:
        d = json.loads(ln)
    except Exception:
        continue
if d is None:
    print(txt[-600:]); sys.exit(1)
r = d.get("roofline") or {}
print("value %.3e ms/step %.4f frac %.3f iso %.4f fill %.4f launches %s host_us %.1f n_gpus %s" % (
    d["value"], d["ms_per_step"], r.get("frac", 0), r.get("isolated_launch_ms", 0), (r.get("fill_phase") or {}).get("kernel_ms", 0),
    d.get("gpu_launches"), d.get("host_enqueue_us_per_step", 0), d.get("n_gpus")))
print("clocks", d.get("clocks"))
for v in d.get("variants") or []:
    print("  %-62s ms %.4f frames/s %.3e frac %.3f launches %.1f items %s" % (v["name"][:62], v["ms_per_step"], v["value"], v["roofline"]["frac"], v["launches_per_step"], v["items"]))
if d.get("reference_api_call"): print("api ms/call %.2f" % d["reference_api_call"]["ms_per_call"])
if d.get("e2e"): print("e2e %.3e ms %.2f copy-only %.2f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("copy_only_ms_per_step", 0)))
if d.get("cpu_baseline"): print("cpu", json.dumps(d["cpu_baseline"])[:400])
if d.get("final_gather_verified") is not None: print("final gather verified:", d["final_gather_verified"])
