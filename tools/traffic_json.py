"""Refresh profiles/<R>_traffic.json from the ncu summaries profiles/<R>_<capture>_ncu.txt (tools/ncu_summary.py output):
the measured fields (dram_read, dram_write, time_us) are replaced, the descriptions and algorithmic byte counts are kept.
  python tools/traffic_json.py r02"""
import json
import os
import re
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
path = os.path.join(root, "%s_traffic.json" % R)
d = json.load(open(path))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def metric(txt, name):
    m = re.search(re.escape(name) + r" \('([0-9.]+)', '([A-Za-z]+)'\)", txt)
    return (float(m.group(1)), m.group(2)) if m else None


for cap, rec in d["captures"].items():
    f = os.path.join(root, "%s_%s_ncu.txt" % (R, cap))
    if not os.path.exists(f):
        continue
    txt = open(f).read()
    r, w, t = metric(txt, "dram__bytes_read.sum"), metric(txt, "dram__bytes_write.sum"), metric(txt, "gpu__time_duration.sum")
    if r and w and t:
        rec["dram_read"] = int(r[0] * UNIT[r[1]])
        rec["dram_write"] = int(w[0] * UNIT[w[1]])
        rec["time_us"] = t[0] if t[1] == "us" else t[0] * 1e3
json.dump(d, open(path, "w"), indent=1)
print(json.dumps({k: (v["dram_read"] + v["dram_write"], v["time_us"]) for k, v in d["captures"].items()}))
