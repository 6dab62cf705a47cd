"""The oracle against itself and against every pin the reference offers for this path:
two independent restatements (numpy-einsum fp64 vs torch-conv fp64), the structural known answers of
train.py:194 / doc/old_readme.md:154, the transposed-conv identity, and the committed golden vectors."""
import os

import numpy as np
import torch

from oracle import dctts_numpy as on
from oracle import dctts_torch as ot
from oracle.params import HP, init_params, ssrn_specs, synthetic_batch, text2mel_specs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_variable_inventory_known_answers():
    hp = HP()
    specs = text2mel_specs(hp)
    d = {n: s for n, s, _ in specs}
    # train.py:194 prints these two for the LJ config (vocab size there was 61; shapes otherwise identical)
    assert d["Text2Mel/TextEnc/embed_1/lookup_table"] == (len(hp.vocab), 128)
    assert d["Text2Mel/TextEnc/C_2/conv1d/kernel"] == (1, 128, 512)
    assert len(specs) == 209
    n = sum(int(np.prod(s)) for _, s, _ in specs)
    assert n == 23974512                      # x4 bytes x3 (w, Adam m, v) = 288 MB ~ "c.300MB" (doc/old_readme.md:154)
    assert sum(int(np.prod(s)) for _, s, _ in ssrn_specs(HP(full_dim=513))) == 25517583
    assert sum(int(np.prod(s)) for _, s, _ in ssrn_specs(HP(full_dim=1025))) == 28410383
    assert len(ssrn_specs(hp)) == 80


def test_numpy_and_torch_restatements_agree():
    hp = HP(max_N=30, max_T=50)
    P = init_params(text2mel_specs(hp), 0, perturb=True)
    b = synthetic_batch(hp, 2, 30, 50, text_len=22)
    a = on.text2mel_forward(hp, P, b["L"], b["mels"], "train")
    t = ot.text2mel_forward(hp, ot.to_torch(P, torch.float64), b["L"], torch.tensor(b["mels"], dtype=torch.float64),
                            "generate_attention")
    for k in ("K", "V", "Q", "R", "alignments", "Y_logits", "Y"):
        assert np.abs(a[k] - t[k].numpy()).max() < 1e-11, k
    assert (a["max_attentions"] == t["max_attentions"].numpy()).all()
    la = on.text2mel_loss(hp, a, b["mels"])
    lt = ot.text2mel_loss(hp, t, torch.tensor(b["mels"], dtype=torch.float64))
    np.testing.assert_allclose(la, [float(x) for x in lt], rtol=1e-12)
    prev = np.array([2, 11])
    a = on.text2mel_forward(hp, P, b["L"], b["mels"], "synthesize", prev)
    t = ot.text2mel_forward(hp, ot.to_torch(P, torch.float64), b["L"], torch.tensor(b["mels"], dtype=torch.float64),
                            "synthesize", prev)
    assert np.abs(a["Y"] - t["Y"].numpy()).max() < 1e-11
    ali = a["alignments"]
    assert (ali[0, :2] == 0).all() and (ali[0, 5:] == 0).all() and (ali[1, :11] == 0).all() and (ali[1, 14:] == 0).all()


def test_ssrn_restatements_agree_and_upsample_by_r():
    hp = HP(full_dim=513)
    P = init_params(ssrn_specs(hp), 1, perturb=True)
    y = synthetic_batch(hp, 2, 8, 12)["mels"]
    lg, Z = on.SSRN(hp, P, y.astype(np.float64))
    lgt, Zt = ot.SSRN(hp, ot.to_torch(P, torch.float64), torch.tensor(y, dtype=torch.float64))
    assert Z.shape == (2, 48, 513)
    assert np.abs(Z - Zt.numpy()).max() < 1e-11


def test_transposed_conv_is_gradient_of_stride2_same_conv():
    """modules.py:243: conv2d_transpose == d/dy of the stride-2 SAME conv (pad_left 0, pad_right 1)."""
    rng = np.random.default_rng(0)
    C, L = 5, 7
    W = rng.standard_normal((1, 3, C, C))
    P = {"d/conv2d_transpose/kernel": W, "d/conv2d_transpose/bias": np.zeros(C),
         "d/normalize/gamma": np.ones(C), "d/normalize/beta": np.zeros(C)}
    x = rng.standard_normal((1, L, C))
    out = on.conv1d_transpose(P, x, "d", normtype=None)
    y = torch.zeros(1, 2 * L, C, dtype=torch.float64, requires_grad=True)
    yp = torch.nn.functional.pad(y.transpose(1, 2), (0, 1))
    k = torch.tensor(W[0]).permute(2, 1, 0)            # forward kernel [k, in=Cout.., out] -> conv1d weight [Cin_fwd_out, Cin_fwd_in, k]
    fwd = torch.nn.functional.conv1d(yp, k, stride=2)  # [1, C(in of transpose), L]
    (fwd * torch.tensor(x).transpose(1, 2)).sum().backward()
    assert np.abs(out - y.grad.numpy()).max() < 1e-12


def test_guide_lr_adam_known_answers():
    W = on.get_attention_guide(4, 5, 0.2)
    assert W.dtype == np.float32 and W[0, 0] == 0.0
    np.testing.assert_allclose(W[1, 0], 1 - np.exp(-(0.25) ** 2 / 0.08), rtol=1e-6)
    np.testing.assert_allclose(on.learning_rate_decay(0.001, 0), 0.001 * 4000 ** 0.5 * 4000 ** -1.5)
    np.testing.assert_allclose(on.learning_rate_decay(0.001, 3999), 0.001, rtol=1e-12)
    p, m, v = on.adam_step(np.array([1.0]), np.zeros(1), np.zeros(1), np.array([5.0]), 1, 0.1)
    # clipped grad 1 -> m=.1 v=.001; lr_t = .1*sqrt(.001)/.1; update = lr_t*.1/(sqrt(.001)+1e-8) ~ 0.1
    np.testing.assert_allclose(p, 1.0 - 0.1 * np.sqrt(0.001) / 0.1 * 0.1 / (np.sqrt(0.001) + 1e-8))


def test_torch_adam_matches_numpy_adam():
    hp = HP(decay_lr=True)
    P = {"w": np.linspace(-1, 1, 7).astype(np.float32)}
    Pt = ot.to_torch(P, torch.float64)
    opt = ot.TFAdam(hp, Pt)
    p, m, v = P["w"].astype(np.float64), np.zeros(7), np.zeros(7)
    for t in range(1, 5):
        g = np.sin(np.arange(7) * t) * 3
        opt.step({"w": torch.tensor(g)})
        p, m, v = on.adam_step(p, m, v, g, t, on.learning_rate_decay(hp.lr, t - 1))
    np.testing.assert_allclose(Pt["w"].numpy(), p, rtol=1e-12)


def test_golden_vectors_reproduce():
    z = np.load(os.path.join(GOLD, "t2m_c1.npz"))
    hp = HP(max_N=60, max_T=200)
    P = init_params(text2mel_specs(hp), int(z["param_seed"]), perturb=True)
    b = synthetic_batch(hp, 2, 60, 200, seed=int(z["data_seed"]), text_len=50)
    t = ot.text2mel_forward(hp, ot.to_torch(P, torch.float32), b["L"], torch.tensor(b["mels"]), "generate_attention")
    assert np.abs(t["Y"].numpy() - z["Y"]).max() < 2e-5          # fp32 restatement vs stored fp64 result
    assert np.abs(t["alignments"].numpy() - z["alignments"]).max() < 1e-5
    comps = ot.text2mel_loss(hp, t, torch.tensor(b["mels"]))
    np.testing.assert_allclose([float(c) for c in comps], z["loss_components"], rtol=1e-4)
    z2 = np.load(os.path.join(GOLD, "ssrn_small.npz"))
    hp2 = HP(full_dim=513)
    Ps = init_params(ssrn_specs(hp2), int(z2["param_seed"]), perturb=True)
    b2 = synthetic_batch(hp2, 2, 8, 24, seed=int(z2["data_seed"]), with_mags=True)
    _, Z = ot.SSRN(hp2, ot.to_torch(Ps, torch.float32), torch.tensor(b2["mels"]))
    assert np.abs(Z.numpy() - z2["Z"]).max() < 2e-5


def test_autoregressive_loop_early_stop():
    hp = HP(max_N=12, max_T=9)
    P = init_params(text2mel_specs(hp), 3, perturb=True)
    L = synthetic_batch(hp, 2, 12, 9, text_len=2)["L"]
    K, V = on.TextEnc(hp, P, L)
    Y, t_ends, ali = on.synth_codedtext2mel(hp, P, K, V, np.array([2, 2]))
    assert Y.shape == (2, 9, 80) and ali.shape == (2, 12, 9) and len(t_ends) == 2
    assert all(0 <= t <= 9 for t in t_ends)
    # the window [prev, prev+3) moves at most 2 positions per frame and never backwards
    am = ali.argmax(1)
    for b in range(2):
        last = t_ends[b] if t_ends[b] < 9 else 8
        steps = np.diff(am[b, :last + 1])
        assert (steps >= 0).all() and (steps <= 2).all()


def test_confidence_terms_restatements_agree_and_match_the_diagnostic_script():
    """architectures.py:283-321 (training losses) in both oracles, and -- for one utterance -- against the independent
    restatement of calculate_CDP_Ain_Aout.py (the reference's own numpy diagnostic of the same quantities)."""
    import torch
    from helpers import make_hp, oracle_params
    from ophelia_b200 import synthesize as syn
    from oracle import dctts_numpy as on
    from oracle import dctts_torch as ot
    from oracle.params import synthetic_batch
    hp = make_hp(max_N=24, max_T=40, dropout_rate=0.0, lw_cdp=0.2, lw_ain=0.3, lw_aout=0.1)
    P = oracle_params(hp, "t2m", seed=1)
    b = synthetic_batch(hp, 2, 20, 36)
    out = on.text2mel_forward(hp, P, b["L"], b["mels"], "train")
    c1 = on.text2mel_loss(hp, out, b["mels"])
    Pt = ot.to_torch(P, torch.float64)
    o2 = ot.text2mel_forward(hp, Pt, torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64), "train")
    c2 = [float(c) for c in ot.text2mel_loss(hp, o2, torch.tensor(b["mels"], dtype=torch.float64))]
    assert len(c1) == len(c2) == 8
    np.testing.assert_allclose(c1, c2, rtol=1e-12)
    assert abs(c1[0] - (hp.lw_mel * c1[1] + hp.lw_bd1 * c1[2] + hp.lw_att * c1[3] + 0.2 * c1[5] + 0.3 * c1[6] + 0.1 * c1[7])) < 1e-12
    for i in range(2):
        A = out["alignments"][i:i + 1]
        cdp, ain, aout = on.attention_confidence_terms(A)
        apin, apout = syn.getAP(A[0])
        assert abs(cdp - syn.getCDP(A[0])) < 1e-12 and abs(ain - apin) < 1e-12 and abs(aout - apout) < 1e-12
    hp.loss_weights = {"t2m": {"L1": 0.3, "binary_divergence": 0.3, "attention": 0.3, "L2": 0.1}}
    c3 = on.text2mel_loss(hp, out, b["mels"])
    assert len(c3) == 8 and abs(c3[0] - (0.3 * c3[1] + 0.3 * c3[2] + 0.3 * c3[3] + 0.1 * c3[4])) < 1e-12


def test_torch_autoregressive_loop_equals_numpy_loop_and_truncation_is_exact():
    """oracle.dctts_torch.synth_codedtext2mel (the loop the GPU routes are checked against at BASELINE's synthesis shape)
    equals the numpy restatement of synthesize.py:150-230, and feeding rows 0..j only (truncate=True) is the same function
    as feeding all max_T rows: the networks are causal and every other op acts per row.  With the synthetic diagonal bias
    on the keys the attention advances and sentences end at different frames."""
    hp = HP(max_N=26, max_T=30)
    P = init_params(text2mel_specs(hp), 3, perturb=True)
    b = synthetic_batch(hp, 3, 26, 30, text_len=9)
    K, V = on.TextEnc(hp, P, b["L"])
    K = ot.advancing_keys(K, c=8.0, seed=1)
    ends = np.array([6, 9, 27])
    Yn, tn, an = on.synth_codedtext2mel(hp, P, K, V, ends)
    Pt = ot.to_torch(P, torch.float64)
    Y1, t1, a1, margin = ot.synth_codedtext2mel(hp, Pt, K, V, ends, truncate=True, return_margin=True)
    Y2, t2, a2 = ot.synth_codedtext2mel(hp, Pt, K, V, ends, truncate=False)
    assert t1 == t2 == tn and margin > 1e-6
    assert min(tn) < hp.max_T - 1 and max(tn) == hp.max_T, tn      # early ends and a sentence that never ends
    assert np.abs(Y1 - Yn).max() < 1e-10 and np.abs(a1 - an).max() < 1e-10
    assert np.abs(Y1 - Y2).max() < 1e-12 and np.abs(a1 - a2).max() < 1e-12
    path = an[0].argmax(0)
    assert len(set(path.tolist())) > 3                              # the window moved
