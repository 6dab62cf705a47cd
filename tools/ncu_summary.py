"""Summarise an .ncu-rep of the GEMM kernel: key raw metrics, stall reasons, SASS instruction mix, hottest lines.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_ctas*warps*kblocks divisor]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
for k in ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
          'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
          'l1tex__m_l1tex2xbar_write_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
          'launch__registers_per_thread', 'launch__grid_size', 'launch__cluster_size']:
    print(k, d.get(k))
for k in sorted(d):
    if 'smsp__average_warps_issue_stalled' in k and 'per_issue' in k and '.ratio' in k and 'not_issued' not in k:
        try:
            if float(d[k][0]) > 0.15:
                print("  stall", k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), d[k][0])
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, iex, ins = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
byop = collections.Counter()
tot = 0
data = []
for r in rows[2:]:
    try:
        n = int(r[iex])
    except (ValueError, IndexError):
        continue
    s = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip())
    op = s.split()[0].split('.')[0] if s else '?'
    byop[op] += n
    tot += n
    try:
        data.append((int(r[ins]), r))
    except ValueError:
        pass
print("total warp-instr", tot, "/div:", tot / div)
print([(op, round(n / div, 1)) for op, n in byop.most_common(16)])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data.sort(key=lambda x: -x[0])
print("total samples", sum(n for n, _ in data))
for n, r in data[:16]:
    st = sorted(((int(r[i]) if r[i] else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(n, r[isrc][:80], r[iex], st)
