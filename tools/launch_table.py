"""Per-launch table of one eager training step: run with OPH_PROF_DUMP=<file> python bench.py --no-graph --no-sub
--no-cpu-baseline --steps 1 --warmup 3, then `python tools/launch_table.py <file>`.  Every line of the dump is one
launch (tag, CUDA-event microseconds, algorithmic FLOPs or bytes, shape); the table shows the LAST dumped pass."""
import sys

TAGS = {0: "other", 1: "conv_fwd", 2: "dgrad", 3: "wgrad", 4: "attention", 5: "row_fwd", 6: "row_bwd", 7: "hc_fwd",
        8: "hc_row_fwd", 9: "ar_enc"}
GEMM = {1, 2, 3, 4, 7}


def main(path):
    passes, cur = [], None
    for line in open(path):
        if line.startswith("# begin"):
            cur = []
            passes.append(cur)
            continue
        f = line.split(None, 3)
        if cur is None or len(f) < 3:
            continue
        cur.append((int(f[0]), float(f[1]), float(f[2]), f[3].strip() if len(f) > 3 else ""))
    rows = passes[-1]
    tot = {}
    print("%-11s %9s %10s  %s" % ("tag", "us", "TF/s|GB/s", "shape"))
    for tag, us, work, desc in rows:
        rate = work / us * (1e-6 if tag in GEMM else 1e-3) if us > 0 else 0.0
        print("%-11s %9.1f %10.1f  %s" % (TAGS.get(tag, tag), us, rate, desc))
        a = tot.setdefault(tag, [0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += work
    print()
    for tag, (n, us, work) in sorted(tot.items()):
        rate = work / us * (1e-6 if tag in GEMM else 1e-3)
        print("%-11s %4d launches %9.1f us  %8.1f %s" % (TAGS.get(tag, tag), n, us, rate, "TFLOP/s" if tag in GEMM else "GB/s"))


if __name__ == "__main__":
    main(sys.argv[1])
