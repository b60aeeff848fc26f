# A/B of ring-mode knobs on a B200: prints ms/step for each "<env assignments>" argument (bench.py --no-e2e --no-cpu)
for cfg in "$@"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 600 2>/dev/null | CFG="$cfg" python -c '
import sys,json,os
d=json.loads(sys.stdin.readlines()[-1]); r=d["roofline"]
print("%-44s ms/step %.4f  dense %.4f  in-flight %.4f  isolated %.4f" % (os.environ["CFG"], d["ms_per_step"], r["kernel_ms"], r["kernel_ms_in_flight"], d["workload_stats"]["isolated_step_ms"]))'
done
