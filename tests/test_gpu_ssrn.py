"""Parity of the CUDA SSRN path (conv1d / hc / conv1d_transpose stack) against the fp64 oracles."""
import os

import numpy as np
import pytest
import torch

from helpers import check_grads, make_hp, maxabs, oracle_params
from oracle import dctts_numpy as on
from oracle import dctts_torch as ot
from oracle.params import synthetic_batch

pytestmark = pytest.mark.gpu


def _graph(hp, mode, P, **kw):
    from ophelia_b200.architectures import SSRNGraph
    from ophelia_b200.variables import VariableStore
    store = VariableStore("cuda:0")
    g = SSRNGraph(hp, mode=mode, store=store, **kw)
    store.load_state_dict(P)
    return g


@pytest.mark.parametrize("full_dim,T", [(513, 24), (1025, 37)])
def test_forward_matches_oracle(full_dim, T):
    from ophelia_b200.session import Session
    hp = make_hp(full_dim=full_dim)
    P = oracle_params(hp, "ssrn", seed=5)
    b = synthetic_batch(hp, 2, 8, T, seed=11)
    logits, Z = on.SSRN(hp, P, b["mels"].astype(np.float64))
    g = _graph(hp, "synthesize", P)
    Zg = Session().run(g.Z, {g.mels: b["mels"]})
    assert Zg.shape == (2, 4 * T, full_dim)
    assert maxabs(Zg, Z) < 1e-3


@pytest.mark.parametrize("full_dim", [513, 1025])
def test_forward_matches_oracle_at_baseline_length(full_dim):
    """BASELINE.json configs[2] length (T=870 -> 3480 frames) with B=2 against the fp64 torch oracle: 6960 output rows,
    so the 1024-channel highway layers (HC_11/12) and the odd-width output layers run several rounds of work units per
    launch (networks.py:437-537)."""
    from ophelia_b200.session import Session
    T = 870
    hp = make_hp(full_dim=full_dim)
    P = oracle_params(hp, "ssrn", seed=8)
    b = synthetic_batch(hp, 2, 8, T, seed=13)
    with torch.no_grad():
        logits, Z = ot.SSRN(hp, ot.to_torch(P, torch.float64), torch.tensor(b["mels"], dtype=torch.float64))
    g = _graph(hp, "synthesize", P)
    Zg, Lg = Session().run([g.Z, g.Z_logits], {g.mels: b["mels"]})
    assert Zg.shape == (2, 4 * T, full_dim)
    assert maxabs(Zg, Z.numpy()) < 1e-3, maxabs(Zg, Z.numpy())
    assert maxabs(Lg, logits.numpy()) < 5e-3, maxabs(Lg, logits.numpy())


def test_forward_matches_golden_fixture():
    from ophelia_b200.session import Session
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ssrn_small.npz"))
    hp = make_hp(full_dim=513)
    P = oracle_params(hp, "ssrn", seed=int(z["param_seed"]))
    b = synthetic_batch(hp, 2, 8, 24, seed=int(z["data_seed"]), with_mags=True)
    g = _graph(hp, "synthesize", P)
    Zg = Session().run(g.Z, {g.mels: b["mels"]})
    assert maxabs(Zg, z["Z"]) < 1e-3


@pytest.mark.parametrize("l1", [True, False])
def test_train_step_matches_oracle(l1):
    hp = make_hp(full_dim=513, dropout_rate=0.0)
    if not l1:
        hp.lw_mag, hp.lw_bd2, hp.lw_ssrn_l2 = 0.0, 0.7, 0.3
    P = oracle_params(hp, "ssrn", seed=6)
    b = synthetic_batch(hp, 2, 8, 33, seed=12, with_mags=True)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    mels = torch.tensor(b["mels"], dtype=torch.float64)
    mags = torch.tensor(b["mags"], dtype=torch.float64)
    g = _graph(hp, "train", P, data=iter([]))
    md, gd = torch.tensor(b["mels"]).cuda(), torch.tensor(b["mags"]).cuda()
    for step in range(2):
        comps_ref, grads_ref = ot.ssrn_train_step(hp, Pt, opt, mels, mags)
        comps = g.train_step_device(md, gd).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=2e-4, atol=1e-6)
        if step == 0:
            check_grads({n: g.store.grads[n].cpu().numpy() for n in grads_ref}, grads_ref, "SSRN/C_16/",
                        strict_tol=1e-2 if l1 else 2e-4)
