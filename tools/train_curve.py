"""Sanity run of the training loop through Session.run on synthetic LJ-shape batches: prints the loss components every
50 steps (train.py:270-309 call pattern).  python tools/train_curve.py [steps]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
from ophelia_b200.architectures import Text2MelGraph
from ophelia_b200.configuration import default_hparams
from ophelia_b200.data import SyntheticBatches
from ophelia_b200.session import Session
from ophelia_b200.variables import VariableStore

__graft_entry__.build()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device("cuda:0")
hp = default_hparams(max_N=180, max_T=870, seed=0)
src = SyntheticBatches(hp, "t2m", 32, N=180, T=870, seed=1234, n_distinct=4)
g = Text2MelGraph(hp, mode="train", store=VariableStore(dev, seed=0), data=src, device=dev)
sess = Session()
t0 = time.perf_counter()
for i in range(steps):
    gs, comps, _ = sess.run([g.global_step, g.loss_components, g.train_op])
    if gs == 1 or gs % 50 == 0:
        print("step %4d  loss %.5f  L1 %.5f  BD %.5f  att %.6f  L2 %.5f  (%.1f s)" % ((gs,) + tuple(comps) + (time.perf_counter() - t0,)), flush=True)
