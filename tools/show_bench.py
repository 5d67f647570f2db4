"""Print the step time, whole-pass roofline and per-kernel milliseconds of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d.get("roofline", {}).get("kernel_ms_per_step", {})
print(round(d["ms_per_step"], 3), "ms/step", "%.4g" % d["value"], d.get("pass_roofline", {}).get("frac"),
      {n: v for n, v in k.items() if v >= 0.25})
