"""Time (CUDA events) or expose to ncu one layer-shaped call of the hot path.
  python tools/perf_layer.py --op hc_fwd --B 32 --L 870 --C 256 --k 3 --iters 20
ops: hc_fwd, hc_bwd, conv_fwd (k=1, C->C), attn_fwd"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ophelia_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="hc_fwd")
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--L", type=int, default=870)
ap.add_argument("--C", type=int, default=256)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--rate", type=int, default=3)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cluster", type=int, default=4)
ap.add_argument("--planes", type=int, default=1)
ap.add_argument("--dbg", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, L, C, k = a.B, a.L, a.C, a.k
x = torch.randn(B, L, C, device=dev)
if a.planes:    # like inside a network: the previous layer's row-wise kernel attached split-bf16 planes to x
    w0 = torch.randn(1, C, C, device=dev) * (2.6 / C) ** 0.5
    x, _, _ = ops.conv1d_fwd(x, ops.PackedConv(w0), torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev))
w = torch.randn(k, C, 2 * C, device=dev) * (2.6 / (k * C)) ** 0.5
pk = ops.PackedConv(w)
bias = torch.zeros(2 * C, device=dev)
g1 = torch.ones(C, device=dev); b1 = torch.zeros(C, device=dev); g2 = torch.ones(C, device=dev); b2 = torch.zeros(C, device=dev)
y = torch.empty_like(x)
dy = torch.randn_like(x)
grads = [torch.zeros_like(t) for t in (w, bias, g1, b1, g2, b2)]
_, saved = ops.hc_fwd(x, pk, bias, g1, b1, g2, b2, a.rate, 1, True, save=True, y=y)
if a.op == "conv_fwd":
    w1 = torch.randn(1, C, C, device=dev) * (2.6 / C) ** 0.5
    pk1 = ops.PackedConv(w1)


def run():
    if a.op == "hc_fwd":
        ops.hc_fwd(x, pk, bias, g1, b1, g2, b2, a.rate, 1, True, y=y)
    elif a.op == "hc_bwd":
        ops.hc_bwd(dy, x, saved, pk, g1, b1, g2, b2, *grads, a.rate, 1, True)
    elif a.op == "hc_dgrad":
        ops.hc_bwd(dy, x, saved, pk, g1, b1, g2, b2, None, *grads[1:], a.rate, 1, True)
    elif a.op == "conv_fwd":
        ops.conv1d_fwd(x, pk1, bias[:C], g1, b1, 1, 1, 0, 0, True, y=y)
    elif a.op == "attn_fwd":
        Q = x[:, :, :256]; K = torch.randn(B, 180, 256, device=dev); V = torch.randn(B, 180, 256, device=dev)
        ops.attention_fwd(Q, K, V)


dbg = torch.zeros(74, 24, dtype=torch.int64, device=dev)
_lib.call("oph_gemm_debug_buffer", dbg.data_ptr())
_lib.set_debug_flags(a.dbg)
for _ in range(a.warmup):
    run()
torch.cuda.synchronize()
lib = _lib.load()
import ctypes
lib.oph_profile_begin()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
prof = (ctypes.c_double * 30)()
lib.oph_profile_end(prof)
ms = e0.elapsed_time(e1) / a.iters
print("%s B%d L%d C%d k%d planes%d dbg%d: %.3f ms per call" % (a.op, B, L, C, k, a.planes, a.dbg, ms))
for i, n in ((0, "other"), (1, "conv_fwd"), (2, "dgrad"), (3, "wgrad"), (4, "attention"), (7, "hc_fwd")):
    if prof[3 * i] > 0:
        print("   gemm[%s]: %d launches/call, %.3f ms each, %.1f TFLOP/s algorithmic" %
              (n, prof[3 * i] / a.iters, prof[3 * i + 1] / prof[3 * i], prof[3 * i + 2] / prof[3 * i + 1] / 1e9))

for i, n in ((5, "row fwd"), (6, "row bwd"), (8, "hc row fwd")):
    if prof[3 * i] > 0:
        print("   %s: %d launches/call, %.3f ms each, %.0f GB/s algorithmic" %
              (n, prof[3 * i] / a.iters, prof[3 * i + 1] / prof[3 * i], prof[3 * i + 2] / prof[3 * i + 1] / 1e6))
d = dbg.cpu().double()
d = d[d[:, 0] > 0]
if len(d):
    print("   last GEMM, per CTA pair (mean cycles): total %.0f  wait_acc %.0f  wait_A %.0f  wait_B %.0f  k-blocks %.0f  -> %.0f cyc/k-block (MMA needs 1536)" %
          (d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), d[:, 4].mean(), (d[:, 0] / d[:, 4]).mean()))
    sys.stdout.flush(); print("   ns from kernel entry (mean / max over pairs): MMA loop end %.0f / %.0f, producers done %.0f / %.0f, after final cluster barrier %.0f / %.0f; clock %.2f GHz" %
          (d[:, 5].mean(), d[:, 5].max(), d[:, 6].mean(), d[:, 6].max(), d[:, 7].mean(), d[:, 7].max(), (d[:, 0] / d[:, 5]).mean()))
