"""Where does the end-to-end (Session.run) step spend host time?  python tools/e2e_probe.py [ssrn [full_dim]]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
from ophelia_b200.architectures import SSRNGraph, Text2MelGraph
from ophelia_b200.configuration import default_hparams
from ophelia_b200.data import SyntheticBatches
from ophelia_b200.session import Session
from ophelia_b200.variables import VariableStore

__graft_entry__.build()
dev = torch.device("cuda:0")
ssrn = len(sys.argv) > 1 and sys.argv[1] == "ssrn"
hp = default_hparams(max_N=180, max_T=870, seed=0, full_dim=int(sys.argv[2]) if len(sys.argv) > 2 else 513)
src = SyntheticBatches(hp, "ssrn" if ssrn else "t2m", 32, N=180, T=870, seed=1234)
g = (SSRNGraph if ssrn else Text2MelGraph)(hp, mode="train", store=VariableStore(dev, seed=0), data=src, device=dev)
fields = (("mel", torch.float32), ("mag", torch.float32)) if ssrn else (("text", torch.int32), ("mel", torch.float32))
sess = Session()
for _ in range(6):
    sess.run([g.global_step, g.loss_components, g.train_op])
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
n = 20
t_all = time.perf_counter()
for _ in range(n):
    t0 = time.perf_counter(); ins = g._next_inputs(fields); tick("next_inputs", t0)
    t0 = time.perf_counter(); comps = g._step_maybe_graphed(*ins); tick("launch", t0)
    t0 = time.perf_counter(); c = comps.cpu(); tick("wait+d2h", t0)
    t0 = time.perf_counter(); gs = int(g.store.global_step.item()); tick("gs.item", t0)
tot = (time.perf_counter() - t_all) / n
print("e2e probe: %.3f ms/step; " % (tot * 1e3) + ", ".join("%s %.3f ms" % (k, v / n * 1e3) for k, v in T.items()))
