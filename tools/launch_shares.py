"""Per-kernel share of one training step from an ncu launch list (gpu__time_duration.sum CSV).
  python tools/launch_shares.py gpurun_out/launches_r01.csv > profiles/r01_step_kernel_shares.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
data = [(int(r[iid]), r[ik], float(r[iv].replace(',', ''))) for r in rows[1:] if r[iid].isdigit()]
ad = [i for i, (_a, k, _v) in enumerate(data) if 'adam_clip' in k]
step = data[ad[-2] + 1:ad[-1] + 1]                      # one whole step: after one optimiser launch up to the next
agg = collections.defaultdict(lambda: [0, 0.0])
for _a, k, v in step:
    name = k.split('(')[0].replace('void ', '').replace('oph::', '').replace('<unnamed>::', '')
    agg[name][0] += 1
    agg[name][1] += v / 1000
tot = sum(v for _n, v in agg.values())
print("one training step under ncu (cold caches, serialised, single stream): %d launches, %.2f ms of kernel time" % (len(step), tot / 1000))
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-58s %4d launches %9.1f us %5.1f%%" % (k[:58], n, v, 100 * v / tot))
