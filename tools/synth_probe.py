"""Timing of the autoregressive device route: python tools/synth_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
from ophelia_b200 import synthesize as syn
from ophelia_b200.architectures import Text2MelGraph
from ophelia_b200.configuration import default_hparams
from ophelia_b200.session import Session
from ophelia_b200.variables import VariableStore
__graft_entry__.build()
dev = torch.device("cuda", 0)
hp = default_hparams(max_N=150, max_T=200, full_dim=513, seed=0)
rng = np.random.default_rng(1234)
L = np.zeros((10, 150), np.int32)
for i in range(10):
    n = int(rng.integers(60, 149)); L[i, :n] = rng.integers(1, len(hp.vocab), n)
g1 = Text2MelGraph(hp, mode="synthesize", store=VariableStore(dev, seed=0), device=dev)
sess = Session()
ends = syn.get_text_lengths(L)
K, V = syn.encode_text(hp, L, g1, sess)
for kw in (dict(use_cuda_graph=True, check_every=8), dict(use_cuda_graph=True, check_every=8), dict(use_cuda_graph=True, check_every=1),
           dict(use_cuda_graph=True, check_every=200), dict(use_cuda_graph=False, check_every=8)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Y, t_ends, _ = syn.synth_codedtext2mel_device(hp, K, V, ends, g1, **kw)
    torch.cuda.synchronize(); print(kw, "%.4f s" % (time.perf_counter() - t0), "frames", sum(t_ends))
# one forward on the device, timed with events
Yd = torch.zeros(10, 200, 80, device=dev); prev = torch.zeros(10, dtype=torch.int32, device=dev)
Kd, Vd = torch.tensor(K).to(dev), torch.tensor(V).to(dev)
from ophelia_b200 import ops
ops.ensure_planes(Kd, cache=True); ops.ensure_planes(Vd, cache=True)
for _ in range(3):
    g1.build_model(None, Yd, False, K=Kd, V=Vd, prev_max_attentions=prev, want_alignments=True)
gr = torch.cuda.CUDAGraph()
torch.cuda.synchronize()
with torch.cuda.graph(gr):
    out = g1.build_model(None, Yd, False, K=Kd, V=Vd, prev_max_attentions=prev, want_alignments=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    gr.replay()
e1.record(); torch.cuda.synchronize()
print("graph replay of one frame step: %.3f ms" % (e0.elapsed_time(e1) / 50))
