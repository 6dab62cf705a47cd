"""SASS mnemonic counts per kernel of libophelia_sm100.so (the evidence that the hot kernels are tcgen05 / TMA code):
  python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ophelia_b200", "libophelia_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = re.compile(r"\b(UTCHMMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UCGABAR_\w+|LDGSTS[.\w]*|REDG?[.\w]*|SYNCS[.\w]*)")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = counts.setdefault(name, collections.Counter())
        continue
    if kern is None:
        continue
    m = WANT.search(line)
    if m:
        op = m.group(1)
        if op.startswith("SYNCS"):
            op = "SYNCS.* (all mbarrier forms)"
        elif op.startswith("LDGSTS"):
            op = "LDGSTS.* (cp.async)"
        elif op.startswith("RED"):
            op = "RED.* (global reductions)"
        kern[op] += 1
print("SASS mnemonic counts of libophelia_sm100.so (cuobjdump -sass, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3), round 2")
print("tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA tensor copy = UTMALDG, tensor-map prefetch = UTMAPF, tcgen05.commit = UTCBAR,")
print("cluster barrier = UCGABAR, mbarrier ops = SYNCS, cp.async = LDGSTS\n")
for name, c in counts.items():
    if not any(k.startswith(("UTC", "LDTM", "UTMA", "LDGSTS")) for k in c):
        continue
    print(name)
    for op, n in sorted(c.items()):
        print("  %-40s %d" % (op, n))
    print()
