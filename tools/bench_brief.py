"""Run bench.py (arguments passed through) and print the headline numbers of its JSON line."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True).stdout.strip().splitlines()
d = json.loads(out[-1])
r, e = d.get("roofline") or {}, d.get("e2e") or {}
print("value %.0f Mpix/s  step %.4f ms  roofline %.3f (%.1f us)  e2e %s  launches %s" % (
    d["value"], d["ms_per_step"], r.get("frac", float("nan")), 1e3 * r.get("kernel_ms", float("nan")),
    ("%.0f Mpix/s %.2f ms" % (e["value"], e["ms_per_step"])) if e else "-", d.get("gpu_launches")))
