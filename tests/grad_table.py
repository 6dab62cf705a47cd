"""Diagnostic: per-variable gradient error of one Text2Mel training step vs the fp64 oracle, in creation order."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_hp, oracle_params, relerr
from oracle import dctts_torch as ot
from oracle.params import synthetic_batch
from ophelia_b200.architectures import Text2MelGraph
from ophelia_b200.variables import VariableStore

B, N, T = 2, 60, 200
hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
P = oracle_params(hp, "t2m", seed=2)
b = synthetic_batch(hp, B, N, T, ragged=True)
Pt = ot.to_torch(P, torch.float64, requires_grad=True)
opt = ot.TFAdam(hp, Pt)
comps_ref, grads_ref = ot.text2mel_train_step(hp, Pt, opt, torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64))
Pf = ot.to_torch(P, torch.float32, requires_grad=True)
_, grads_f32 = ot.text2mel_train_step(hp, Pf, ot.TFAdam(hp, Pf), torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"]))
store = VariableStore("cuda:0")
g = Text2MelGraph(hp, mode="train", store=store, data=iter([]))
store.load_state_dict(P)
comps = g.train_step_device(torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()).cpu().numpy()
print("loss comps", comps, comps_ref)
for n in grads_ref:
    if n.endswith("kernel") or n.endswith("lookup_table"):
        print("%-50s ours %.2e   torch-fp32 %.2e   |g| %.2e" % (n, relerr(store.grads[n].cpu().numpy(), grads_ref[n].numpy()),
              relerr(grads_f32[n].numpy(), grads_ref[n].numpy()), float(grads_ref[n].norm())))
