for sk in 1 0; do
  ISG_SPLIT_KEEP=$sk python bench.py --no-e2e --no-cpu --steps 600 "$@" 2>/dev/null | SK=$sk python -c '
import sys,json,os
d=json.loads(sys.stdin.readlines()[-1]); r=d["roofline"]
print("split", os.environ["SK"], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(r["kernel_ms"],4), "in_flight", round(r["kernel_ms_in_flight"],4), "iso", d["workload_stats"]["isolated_step_ms"], "frac", round(r["frac"],3))'
done
