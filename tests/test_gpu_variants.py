"""The research variants on the same operator surface against the fp64 torch restatement in
oracle/dctts_variants.py: speaker embeddings at the positions of hp.multispeaker, MerlinTextEnc / LinearTransformLabels /
label-only text encoders, FixedAttention with external durations, BabblerGraph and TextEncGraph."""
import numpy as np
import pytest
import torch

from helpers import check_grads, make_hp, maxabs, network_grad_errors
from oracle import dctts_torch as ot
from oracle import dctts_variants as ov
from oracle.params import init_params, synthetic_batch

pytestmark = pytest.mark.gpu


def _t(x, dt=torch.float64):
    return torch.tensor(np.asarray(x), dtype=dt)


def _params(specs, seed):
    return init_params([(n, s, k) for n, s, k in specs], seed, perturb=True)


def _t2m(hp, mode, P, **kw):
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.variables import VariableStore
    store = VariableStore("cuda:0")
    g = Text2MelGraph(hp, mode=mode, store=store, **kw)
    store.load_state_dict(P)
    return g


def _hard_durations(rng, B, T, N, n_real):
    """Monotonic hard alignments like utils.durations_to_hard_attention_matrix: every frame attends to one symbol."""
    D = np.zeros((B, T, N), np.float32)
    for b in range(B):
        cuts = np.sort(rng.choice(np.arange(1, T), n_real - 1, replace=False))
        sym = np.searchsorted(cuts, np.arange(T), side="right")
        D[b, np.arange(T), sym] = 1.0
    return D


@pytest.mark.parametrize("positions", [
    ['text_encoder_input', 'audio_decoder_input'],
    ['text_encoder_towards_end', 'audio_encoder_input'],
])
def test_multispeaker_text2mel_matches_oracle(positions):
    from ophelia_b200.architectures import text2mel_variables
    from ophelia_b200.session import Session
    B, N, T = 3, 30, 70
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0, multispeaker=positions, nspeakers=5, speaker_embedding_size=32)
    P = _params(text2mel_variables(hp), 21)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    speakers = np.array([[1], [4], [2]], np.int32)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Lt, mt = _t(b["L"], torch.long), _t(b["mels"])
    # forward through the Session surface (g.speakers placeholder, synthesize.py:83-86)
    with torch.no_grad():
        ref = ov.text2mel_forward(hp, Pt, Lt, mt, "generate_attention", speakers=speakers)
    g = _t2m(hp, "generate_attention", P)
    Y, ali = Session().run([g.Y, g.alignments], {g.L: b["L"], g.mels: b["mels"], g.speakers: speakers})
    assert maxabs(Y, ref["Y"].numpy()) < 1e-3 and maxabs(ali, ref["alignments"].numpy()) < 1e-4
    # two optimiser steps; the speaker tables get gradients only in the rows of the speakers in the batch
    gt = _t2m(hp, "train", P, data=iter([]))
    Ld, md, sd = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda(), torch.tensor(speakers).cuda()
    for step in range(2):
        comps_ref, grads_ref = ov.text2mel_train_step(hp, Pt, opt, Lt, mt, speakers=speakers)
        comps = gt.train_step_device(Ld, md, sd).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)
        if step == 0:
            ours = {n: gt.store.grads[n].cpu().numpy() for n in grads_ref}
            check_grads(ours, grads_ref, "Text2Mel/AudioDec/C_", strict_tol=5e-2, median_tol=3e-2, max_tol=1e-1)
            tables = [n for n in grads_ref if "/embed_" in n and grads_ref[n].shape[0] == hp.nspeakers]
            assert len(tables) == len(positions)
            for n in tables:
                gr = ours[n]
                assert np.abs(gr[[0, 3]]).max() == 0.0 and np.abs(gr[[1, 2, 4]]).max() > 0, n


def test_learn_channel_contributions_matches_oracle():
    """'learn_channel_contributions' in hp.multispeaker (modules.py:78-88): a sigmoid gate per (speaker, channel) behind the
    conv1d layers and on the transformation connection of the highway layers of TextEnc / AudioEnc / AudioDec, including the
    mel logits.  Forward, two optimiser steps and the gradients of the gate tables against the fp64 restatement."""
    from ophelia_b200.architectures import text2mel_variables
    from ophelia_b200.session import Session
    B, N, T = 3, 24, 50
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0, multispeaker=['learn_channel_contributions'], nspeakers=4,
                 speaker_embedding_size=16)
    specs = text2mel_variables(hp)
    assert sum("/lcc_embed/" in n for n, _, _ in specs) == 14 + 13 + 10           # TextEnc C,C,12 HC; AudioEnc 3 C, 10 HC; AudioDec 6 HC, 4 C
    P = _params(specs, 27)
    for n in P:                                   # gates away from the trivial 0.5
        if "/lcc_embed/" in n:
            P[n] = (P[n] * 8).astype(np.float32)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    speakers = np.array([[2], [0], [3]], np.int32)         # code 0 = the zero-pad row: gate 0.5, no gradient
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Lt, mt = _t(b["L"], torch.long), _t(b["mels"])
    with torch.no_grad():
        ref = ov.text2mel_forward(hp, Pt, Lt, mt, "generate_attention", speakers=speakers)
    g = _t2m(hp, "generate_attention", P)
    Y, ali = Session().run([g.Y, g.alignments], {g.L: b["L"], g.mels: b["mels"], g.speakers: speakers})
    assert maxabs(Y, ref["Y"].numpy()) < 1e-3 and maxabs(ali, ref["alignments"].numpy()) < 1e-4
    gt = _t2m(hp, "train", P, data=iter([]))
    Ld, md, sd = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda(), torch.tensor(speakers).cuda()
    for step in range(2):
        comps_ref, grads_ref = ov.text2mel_train_step(hp, Pt, opt, Lt, mt, speakers=speakers)
        comps = gt.train_step_device(Ld, md, sd).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)
        if step == 0:
            ours = {n: gt.store.grads[n].cpu().numpy() for n in grads_ref}
            tables = [n for n in grads_ref if "/lcc_embed/" in n]
            assert len(tables) == 37
            for n in tables:
                assert np.abs(ours[n][[0, 1]]).max() == 0.0, n                    # pad row and the absent speaker
            nets = network_grad_errors(ours, {n: grads_ref[n].numpy() for n in tables},
                                       ["Text2Mel/TextEnc/", "Text2Mel/AudioEnc/", "Text2Mel/AudioDec/"])
            assert max(nets.values()) < 5e-2, nets


def test_multispeaker_ssrn_matches_oracle():
    from ophelia_b200.architectures import SSRNGraph, ssrn_variables
    from ophelia_b200.variables import VariableStore
    hp = make_hp(full_dim=513, dropout_rate=0.0, multispeaker=['ssrn_input'], nspeakers=4, speaker_embedding_size=16)
    P = _params(ssrn_variables(hp), 22)
    b = synthetic_batch(hp, 2, 8, 21, seed=5, with_mags=True)
    speakers = np.array([[3], [1]], np.int32)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    store = VariableStore("cuda:0")
    g = SSRNGraph(hp, mode="train", store=store, data=iter([]))
    store.load_state_dict(P)
    md, gd, sd = torch.tensor(b["mels"]).cuda(), torch.tensor(b["mags"]).cuda(), torch.tensor(speakers).cuda()
    for _ in range(2):
        comps_ref, grads_ref = ov.ssrn_train_step(hp, Pt, opt, _t(b["mels"]), _t(b["mags"]), speakers)
        comps = g.train_step_device(md, gd, sd).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)


def test_external_durations_fixed_attention_matches_oracle():
    """hp.use_external_durations (networks.py:327-358): R = durations . V, alignments = durations^T, trivial argmax; K is
    computed but unused, so TextEnc's gradient reaches it through V only."""
    from ophelia_b200.architectures import text2mel_variables
    from ophelia_b200.session import Session
    B, N, T = 2, 28, 66
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0, use_external_durations=True)
    P = _params(text2mel_variables(hp), 23)
    b = synthetic_batch(hp, B, N, T, text_len=20)
    D = _hard_durations(np.random.default_rng(0), B, T, N, 20)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Lt, mt = _t(b["L"], torch.long), _t(b["mels"])
    with torch.no_grad():
        ref = ov.text2mel_forward(hp, Pt, Lt, mt, "synthesize", durations=_t(D))
    g = _t2m(hp, "synthesize", P)
    Y, ali, mx = Session().run([g.Y, g.alignments, g.max_attentions],
                               {g.L: b["L"], g.mels: b["mels"], g.durations: D, g.prev_max_attentions: np.zeros(B, np.int32)})
    assert maxabs(Y, ref["Y"].numpy()) < 1e-3
    assert np.array_equal(ali, D.transpose(0, 2, 1)) and np.array_equal(mx, D.argmax(-1))
    gt = _t2m(hp, "train", P, data=iter([]))
    Ld, md, dd = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda(), torch.tensor(D).cuda()
    for step in range(2):
        comps_ref, grads_ref = ov.text2mel_train_step(hp, Pt, opt, Lt, mt, durations=_t(D))
        comps = gt.train_step_device(Ld, md, dd).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)
        if step == 0:
            ours = {n: gt.store.grads[n].cpu().numpy() for n in grads_ref}
            nets = network_grad_errors(ours, grads_ref, ["Text2Mel/TextEnc/", "Text2Mel/AudioEnc/", "Text2Mel/AudioDec/"])
            assert max(nets.values()) < 5e-2, nets


@pytest.mark.parametrize("enc,phone_emb", [("MerlinTextEnc", False), ("MerlinTextEnc", True), ("minimal_feedforward", False),
                                           ("none", False)])
def test_label_text_encoders_match_oracle(enc, phone_emb):
    """hp.text_encoder_type in MerlinTextEnc / minimal_feedforward / none (architectures.py:192-206, networks.py:15-119,
    540-560): linguistic label vectors per symbol instead of (or next to) the phone embeddings."""
    from ophelia_b200.architectures import text2mel_variables
    B, N, T = 2, 26, 60
    labdim = 256 if enc == "none" else 52
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0, text_encoder_type=enc, merlin_label_dir="labels", merlin_lab_dim=labdim,
                 MerlinTextEncWithPhoneEmbedding=phone_emb)
    P = _params(text2mel_variables(hp), 24)
    b = synthetic_batch(hp, B, N, T, text_len=19)
    labels = np.random.default_rng(1).uniform(0.0, 1.0, (B, N, labdim)).astype(np.float32)
    labels[:, 19:] = 0.0
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Lt, mt = _t(b["L"], torch.long), _t(b["mels"])
    gt = _t2m(hp, "train", P, data=iter([]))
    Ld, md, ld_ = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda(), torch.tensor(labels).cuda()
    for step in range(2):
        comps_ref, grads_ref = ov.text2mel_train_step(hp, Pt, opt, Lt, mt, labels=_t(labels))
        comps = gt.train_step_device(Ld, md, ld_).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)
        if step == 0 and grads_ref:
            ours = {n: gt.store.grads[n].cpu().numpy() for n in grads_ref}
            prefixes = sorted({"/".join(n.split("/")[:2]) + "/" for n in grads_ref})
            nets = network_grad_errors(ours, grads_ref, prefixes)
            assert max(nets.values()) < 5e-2, nets


def test_babbler_and_textenc_graphs_match_oracle():
    from ophelia_b200 import synthesize as syn
    from ophelia_b200.architectures import BabblerGraph, TextEncGraph, text2mel_variables
    from ophelia_b200.session import Session
    from ophelia_b200.variables import VariableStore
    B, T = 3, 24
    hp = make_hp(max_N=20, max_T=T, dropout_rate=0.0)
    hp.loss_weights = {'babbler': {'L1': 0.4, 'binary_divergence': 0.6}}
    hp.batchsize = {'t2m': 4, 'ssrn': 4, 'babbler': 3}
    specs = text2mel_variables(hp, with_text_encoder=False)
    assert all("/TextEnc/" not in n for n, _, _ in specs) and len(specs) == 128
    P = _params(specs, 25)
    b = synthetic_batch(hp, B, 20, T, ragged=True)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    store = VariableStore("cuda:0")
    g = BabblerGraph(hp, mode="train", store=store, data=iter([]))
    store.load_state_dict(P)
    md = torch.tensor(b["mels"]).cuda()
    for _ in range(2):
        comps_ref, _ = ov.babbler_train_step(hp, Pt, opt, _t(b["mels"]))
        comps = g.train_step_device(md).cpu().numpy()
        assert comps.shape == (3,)
        np.testing.assert_allclose(comps, comps_ref, rtol=3e-4, atol=1e-6)
    gs = BabblerGraph(hp, mode="synthesize", store=store)
    Y = syn.synth_babble(hp, gs, Session(), nsamples=2)
    Yr = ov.synth_babble(hp, {k: v.detach() for k, v in Pt.items()}, 2).numpy()
    assert Y.shape == (2, T, 80) and maxabs(Y, Yr) < 1e-3
    # TextEncGraph: K, V only (architectures.py:367-376)
    hp2 = make_hp(max_N=20, max_T=T)
    P2 = _params(text2mel_variables(hp2, with_audio=False), 26)
    st2 = VariableStore("cuda:0")
    ge = TextEncGraph(hp2, mode="synthesize", store=st2)
    st2.load_state_dict(P2)
    K, V = Session().run([ge.K, ge.V], {ge.L: b["L"]})
    with torch.no_grad():
        Kr, Vr = ot.TextEnc(hp2, ot.to_torch(P2, torch.float64), b["L"])
    assert maxabs(K, Kr.numpy()) < 1e-3 and maxabs(V, Vr.numpy()) < 1e-3
