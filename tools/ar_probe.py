"""One eager (graph-free) run of the incremental autoregressive route at BASELINE config 4 (10 sentences, max_N=150,
max_T=200), for a per-kernel launch list:

  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5900 -c 120 --csv \
      --log-file gpurun_out/launches_ar.csv python tools/ar_probe.py --frames 120

(~59 launches per frame step: 26 frame-step kernels for AudioEnc, window gather, Attention and AudioDec over the 85-row
window, scatter, advance).  Without ncu it prints the wall time per frame of the eager and the CUDA-graph runs."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402
from ophelia_b200 import _lib  # noqa: E402
from ophelia_b200 import synthesize as syn  # noqa: E402
from ophelia_b200.architectures import Text2MelGraph  # noqa: E402
from ophelia_b200.configuration import default_hparams  # noqa: E402
from ophelia_b200.session import Session  # noqa: E402
from ophelia_b200.variables import VariableStore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--graph", action="store_true")
args = ap.parse_args()
__graft_entry__.build()
dev = torch.device("cuda", 0)
hp = default_hparams(max_N=150, max_T=args.frames, full_dim=513, seed=0)
rng = np.random.default_rng(1234)
L = np.zeros((10, 150), np.int32)
for i in range(10):
    n = int(rng.integers(60, 149))
    L[i, :n] = rng.integers(1, len(hp.vocab), n)
g1 = Text2MelGraph(hp, mode="synthesize", store=VariableStore(dev, seed=0), device=dev)
K, V = syn.encode_text(hp, L, g1, Session())
ends = np.full(10, hp.max_N + 1)
lib = _lib.load()
results = {}
for use_graph, fused in (((True, False), (True, True), (False, False)) if args.graph else ((False, False),)):
    for rep in range(3 if use_graph else 1):
        n0 = lib.oph_launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Y, t_ends, A = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g1, use_cuda_graph=use_graph, check_every=8,
                                                           fused_encoder=fused)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("cuda_graph=%s fused_encoder=%s run %d: %.3f ms per frame, %d library launches"
              % (use_graph, fused, rep, 1e3 * dt / args.frames, lib.oph_launch_count() - n0), flush=True)
    results[(use_graph, fused)] = (Y, t_ends, A)
if (True, True) in results:
    (Y0, t0_, A0), (Y1, t1_, A1) = results[(True, False)], results[(True, True)]
    print("fused vs per-layer encoder: t_ends equal %s, max |dY| %.2e, max |dA| %.2e"
          % (t0_ == t1_, float(np.abs(Y0 - Y1).max()), float(np.abs(A0 - A1).max())))
