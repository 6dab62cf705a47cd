"""Where a GEMM launch spends its time: in-kernel globaltimer stamps per CTA pair (oph_gemm_debug_buffer, [74][16]).
Runs layer-shaped calls of the hot path and prints, for the LAST GEMM launch of each call, mean / max over the pairs of
  entry skew | prologue end | first operands landed | first accumulator complete | MMA loop end | last epilogue end | exit
(all ns from that pair's kernel entry; skew = entry relative to the first pair; span = last exit - first entry).

    python tools/gemm_phases.py [--carve 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ophelia_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
if a.flags:
    _lib.set_debug_flags(a.flags)
dbg = torch.zeros(74, 24, dtype=torch.int64, device=dev)


def planes(B, L, C):
    """an activation with operand planes attached, like inside a network"""
    x = torch.randn(B, L, C, device=dev)
    w0 = torch.randn(1, C, C, device=dev) * (2.6 / C) ** 0.5
    y, _, _ = ops.conv1d_fwd(x, ops.PackedConv(w0), torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev))
    return y


def report(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us_plain = e0.elapsed_time(e1) / a.iters * 1e3
    _lib.call("oph_gemm_debug_buffer", dbg.data_ptr())
    dbg.zero_()
    fn()
    torch.cuda.synchronize()
    _lib.call("oph_gemm_debug_buffer", None)
    d = dbg.cpu().double()
    d = d[d[:, 8] > 0]
    t0 = d[:, 8].min()
    skew = d[:, 8] - t0
    span = d[:, 13].max() - t0
    cols = (("skew", skew), ("prologue", d[:, 9]), ("1st operands", d[:, 10]), ("1st acc full", d[:, 11]), ("mma end", d[:, 5]),
            ("epi end", d[:, 12]), ("exit", d[:, 7]))
    print("%-34s %6.1f us/call (events, whole call)  last GEMM: %d pairs, span %.1f us, %d k-blocks/pair, %.0f cyc/k-block" %
          (name, us_plain, len(d), span / 1e3, d[:, 4].mean(), (d[:, 0] / d[:, 4].clamp(min=1)).mean()))
    print("      " + "  ".join("%s %.1f/%.1f" % (n, v.mean() / 1e3, v.max() / 1e3) for n, v in cols) + "   (us, mean/max)")


B, L = 32, 870
x256 = planes(B, L, 256)
x512 = planes(B, 180, 512)
one = lambda C: (torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev))  # noqa: E731
for C, x, nm in ((256, x256, "AudioEnc"), (512, x512, "TextEnc")):
    w1 = ops.PackedConv(torch.randn(1, C, C, device=dev) * (2.6 / C) ** 0.5)
    b, g, be = one(C)
    y = torch.empty_like(x)
    report("%s 1x1 conv fwd (GEMM + LN tail)" % nm, lambda: ops.conv1d_fwd(x, w1, b, g, be, 1, 1, 0, 0, True, y=y))
    w3 = ops.PackedConv(torch.randn(3, C, 2 * C, device=dev) * (2.6 / (3 * C)) ** 0.5)
    b2 = torch.zeros(2 * C, device=dev)
    g1, b1 = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    report("%s highway fwd (GEMM + tail)" % nm, lambda: ops.hc_fwd(x, w3, b2, g1, b1, g1, b1, 3, 1, True, y=y))
A = torch.randn(256, 256, device=dev)
Bm = torch.randn(256, 256, device=dev)
report("tiny gemm_nt 256^3", lambda: ops.gemm_nt(A, Bm))
