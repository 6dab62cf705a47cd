"""Print the key numbers of bench.py JSON lines read from stdin (one line per result)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    lm = d.get("launch_mode") or {}
    rf = d.get("roofline") or {}
    print(" ".join(sys.argv[1:]), "n_gpus", d.get("n_gpus"), "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"],
          "eager %.3f" % (lm.get("ms_per_step_eager") or 0), "e2e_ms %.3f" % ((d.get("e2e") or {}).get("ms_per_step") or 0),
          "frac %.4f" % (rf.get("frac") or 0), "launches", d.get("gpu_launches_per_step"), "collective", d.get("collective"),
          "hc_stack_ms", (d.get("hc_stack") or {}).get("ms_forward"))
