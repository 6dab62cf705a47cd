"""Forward + one training step of Text2Mel against the oracles over awkward shapes (tile / k-block / item boundaries).
  python tests/shape_fuzz.py [n_random]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_hp, oracle_params, maxabs
from oracle import dctts_numpy as on
from oracle import dctts_torch as ot
from oracle.params import synthetic_batch
from ophelia_b200.architectures import Text2MelGraph
from ophelia_b200.session import Session
from ophelia_b200.variables import VariableStore

def check(B, N, T):
    """(mel max-abs error, attention max-abs error, relative error of the loss components) for one shape."""
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    P = oracle_params(hp, "t2m", seed=B + N + T)
    b = synthetic_batch(hp, B, N, T, ragged=True, text_len=max(1, N - 2))
    ref = on.text2mel_forward(hp, P, b["L"], b["mels"], "generate_attention")
    store = VariableStore("cuda:0")
    g = Text2MelGraph(hp, mode="generate_attention", store=store)
    store.load_state_dict(P)
    Y, ali = Session().run([g.Y, g.alignments], {g.L: b["L"], g.mels: b["mels"]})
    eY, eA = maxabs(Y, ref["Y"]), maxabs(ali, ref["alignments"])
    # one training step: loss components vs the fp64 torch oracle
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    comps_ref, _ = ot.text2mel_train_step(hp, Pt, ot.TFAdam(hp, Pt), torch.tensor(b["L"].astype(np.int64)),
                                          torch.tensor(b["mels"], dtype=torch.float64))
    st2 = VariableStore("cuda:0")
    gt = Text2MelGraph(hp, mode="train", store=st2, data=iter([]))
    st2.load_state_dict(P)
    comps = gt.train_step_device(torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()).cpu().numpy()
    eL = float(np.abs(comps - np.asarray(comps_ref)).max() / max(1e-9, np.abs(np.asarray(comps_ref)).max()))
    return eY, eA, eL


SHAPES = [(1, 5, 1), (1, 9, 127), (2, 64, 128), (3, 65, 129), (2, 33, 255), (1, 128, 257), (5, 17, 64), (4, 40, 100)]

if __name__ == "__main__":
    shapes = list(SHAPES)
    rng = np.random.default_rng(0)
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        shapes.append((int(rng.integers(1, 5)), int(rng.integers(3, 70)), int(rng.integers(2, 300))))
    bad = 0
    for (B, N, T) in shapes:
        eY, eA, eL = check(B, N, T)
        ok = eY < 1e-3 and eA < 1e-4 and eL < 5e-4
        bad += not ok
        print("B=%d N=%3d T=%3d  mel err %.2e  attention err %.2e  loss rel err %.2e  %s" % (B, N, T, eY, eA, eL, "ok" if ok else "FAIL"), flush=True)
    print("==== %d shapes, %d failed" % (len(shapes), bad))
    sys.exit(1 if bad else 0)
