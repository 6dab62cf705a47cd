"""Debug aid: where do outputs before frame 500 change when only mels after frame 500 change?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_hp, oracle_params
from oracle.params import synthetic_batch
from ophelia_b200.architectures import Text2MelGraph
from ophelia_b200.session import Session
from ophelia_b200.variables import VariableStore
B, N, T = 32, 180, 870
hp = make_hp(max_N=N, max_T=T)
P = oracle_params(hp, "t2m", seed=7)
b = synthetic_batch(hp, B, N, T, ragged=True)
store = VariableStore("cuda:0")
g = Text2MelGraph(hp, mode="generate_attention", store=store)
store.load_state_dict(P)
sess = Session()
f = [g.Y, g.Q, g.R, g.alignments, g.Y_logits]
a = sess.run(f, {g.L: b["L"], g.mels: b["mels"]})
a1 = sess.run(f, {g.L: b["L"], g.mels: b["mels"]})
m2 = b["mels"].copy(); m2[:, 500:] = 1.0 - m2[:, 500:]
c = sess.run(f, {g.L: b["L"], g.mels: m2})
for name, x, x1, y in zip(["Y", "Q", "R", "ali", "logits"], a, a1, c):
    if name == "ali":
        d = np.abs(x[:, :, :501] - y[:, :, :501]); d1 = np.abs(x[:, :, :501] - x1[:, :, :501])
    else:
        d = np.abs(x[:, :501] - y[:, :501]); d1 = np.abs(x[:, :501] - x1[:, :501])
    idx = np.unravel_index(d.argmax(), d.shape)
    print(name, "max diff before 501:", d.max(), "at", idx, "count>0:", int((d > 0).sum()), " run-to-run same input:", d1.max())
