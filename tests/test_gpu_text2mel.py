"""Whole-graph parity of the CUDA Text2Mel path against the fp64 oracles (config C1: B=2, N=60, T=200)."""
import numpy as np
import pytest
import torch

from helpers import argmax_mismatches, check_grads, make_hp, maxabs, network_grad_errors, oracle_params
from oracle import dctts_numpy as on
from oracle import dctts_torch as ot
from oracle.params import synthetic_batch

pytestmark = pytest.mark.gpu


def _graph(hp, mode, P, **kw):
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.variables import VariableStore
    store = kw.pop("store", None) or VariableStore("cuda:0")
    g = Text2MelGraph(hp, mode=mode, store=store, **kw)
    if P is not None:
        store.load_state_dict(P)
    return g


def test_forward_matches_oracle_c1():
    from ophelia_b200.session import Session
    hp = make_hp(max_N=60, max_T=200)
    P = oracle_params(hp, "t2m", seed=0)
    b = synthetic_batch(hp, 2, 60, 200, text_len=50)
    ref = on.text2mel_forward(hp, P, b["L"], b["mels"], "generate_attention")
    g = _graph(hp, "generate_attention", P)
    sess = Session()
    Y, ali, mx, K, V = sess.run([g.Y, g.alignments, g.max_attentions, g.K, g.V], {g.L: b["L"], g.mels: b["mels"]})
    assert Y.shape == (2, 200, 80) and ali.shape == (2, 60, 200) and mx.shape == (2, 200)
    assert maxabs(K, ref["K"]) < 1e-3 and maxabs(V, ref["V"]) < 1e-3
    assert maxabs(Y, ref["Y"]) < 1e-3            # north-star tolerance: mels within 1e-3 max-abs
    assert maxabs(ali, ref["alignments"]) < 1e-4
    # exact, except where the oracle's own top-2 gap is below 1e-5
    bad, ties = argmax_mismatches(mx, ref["max_attentions"], np.swapaxes(ref["alignments"], 1, 2))
    assert bad == 0, (bad, ties)


def test_forward_matches_golden_fixture():
    import os
    from ophelia_b200.session import Session
    path = os.path.join(os.path.dirname(__file__), "golden", "t2m_c1.npz")
    z = np.load(path)
    hp = make_hp(max_N=60, max_T=200)
    P = oracle_params(hp, "t2m", seed=int(z["param_seed"]))
    b = synthetic_batch(hp, 2, 60, 200, seed=int(z["data_seed"]), text_len=50)
    g = _graph(hp, "generate_attention", P)
    Y, ali = Session().run([g.Y, g.alignments], {g.L: b["L"], g.mels: b["mels"]})
    assert maxabs(Y, z["Y"]) < 1e-3
    assert maxabs(ali, z["alignments"]) < 1e-4


def test_synthesis_mode_window_mask():
    from ophelia_b200.session import Session
    hp = make_hp(max_N=60, max_T=200)
    P = oracle_params(hp, "t2m", seed=1)
    b = synthetic_batch(hp, 2, 60, 200, text_len=50)
    prev = np.array([3, 41], np.int32)
    ref = on.text2mel_forward(hp, P, b["L"], b["mels"], "synthesize", prev)
    g = _graph(hp, "synthesize", P)
    sess = Session()
    K, V = sess.run([g.K, g.V], {g.L: b["L"]})
    Y, mx, ali = sess.run([g.Y, g.max_attentions, g.alignments],
                          {g.K: K, g.V: V, g.mels: b["mels"], g.prev_max_attentions: prev})
    assert maxabs(Y, ref["Y"]) < 1e-3
    assert maxabs(ali, ref["alignments"]) < 1e-4
    assert (ali[0, :3] == 0).all() and (ali[0, 6:] == 0).all()       # only keys [prev, prev+3) survive
    bad, ties = argmax_mismatches(mx, ref["max_attentions"], np.swapaxes(ref["alignments"], 1, 2))
    assert bad == 0, (bad, ties)
    # hp.turn_off_monotonic_for_synthesis: no window, keys from each sentence's first padding position + 1 on are masked
    # (networks.py:307-309 with hp.text_lengths set as in synthesize.py:505-507)
    hp.turn_off_monotonic_for_synthesis = True
    hp.text_lengths = np.array([50, 31]) + 1
    ref = on.text2mel_forward(hp, P, b["L"], b["mels"], "synthesize", prev)
    Y, ali = sess.run([g.Y, g.alignments], {g.K: K, g.V: V, g.mels: b["mels"], g.prev_max_attentions: prev})
    assert maxabs(Y, ref["Y"]) < 1e-3 and maxabs(ali, ref["alignments"]) < 1e-4
    assert (ali[0, 51:] == 0).all() and (ali[1, 32:] == 0).all() and (ali[1, :32] > 0).all()


@pytest.mark.parametrize("shape,l1", [((2, 60, 200), True), ((3, 37, 131), True), ((3, 37, 131), False)])
def test_train_step_matches_oracle(shape, l1):
    B, N, T = shape
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    if not l1:      # |Y - mel| has a kink too: without the L1 term the ReLU-free tail must match tightly
        hp.lw_mel, hp.lw_bd1, hp.lw_att, hp.lw_t2m_l2 = 0.0, 0.5, 0.3, 0.2
    P = oracle_params(hp, "t2m", seed=2)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    L = torch.tensor(b["L"].astype(np.int64))
    mels = torch.tensor(b["mels"], dtype=torch.float64)
    g = _graph(hp, "train", P, data=iter([]))
    Ld = torch.tensor(b["L"]).cuda()
    md = torch.tensor(b["mels"]).cuda()
    for step in range(3):
        comps_ref, grads_ref = ot.text2mel_train_step(hp, Pt, opt, L, mels)
        comps = g.train_step_device(Ld, md).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=2e-4, atol=1e-6)
        if step == 0:
            sd = g.store.grads
            check_grads({n: sd[n].cpu().numpy() for n in grads_ref}, grads_ref, "Text2Mel/AudioDec/C_11/",
                            strict_tol=1e-2 if l1 else 2e-4)
            row0 = sd["Text2Mel/TextEnc/embed_1/lookup_table"][0].abs().max().item()
            assert row0 == 0.0                                   # zero-pad row gets no gradient (modules.py:38-40)
    assert int(g.store.global_step.item()) == 3
    new = g.store.state_dict()
    for name, p in Pt.items():
        # Adam's first steps move every weight by ~lr_t regardless of gradient size, so compare absolutely
        assert maxabs(new[name], p.detach().numpy()) < 5e-6 + 0.2 * 3 * ot.noam(hp.lr, 2), name


@pytest.mark.parametrize("fa,gshape", [(False, (3, 30, 80)), (False, (3, 37, 95)), (True, (3, 37, 90)), (True, (3, 31, 77))])
def test_train_step_with_batch_guides_matches_oracle(fa, gshape):
    """hp.attention_guide_dir: the attention targets come with the batch (architectures.py:57-58) -- per-utterance guides
    padded with 1.0 (:263) or, with hp.attention_guide_fa, forced-alignment targets under an MSE loss (:271-280).
    First with the attention term alone (its gradient reaches the encoders through Q and K only: tight bound on the
    ReLU-free AudioEnc tail), then two optimiser steps with all terms."""
    B, N, T = 3, 37, 90
    rng = np.random.default_rng(3)
    gts = rng.uniform(0.0, 1.0, gshape).astype(np.float32)
    gts[1, gshape[1] - 5:, :] = 0.0                                   # dynamic_pad zeros of a shorter utterance
    gts[1, :, gshape[2] - 9:] = 0.0
    b = synthetic_batch(make_hp(), B, N, T, ragged=True)
    L, mels = torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64)
    Ld, md, gd = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda(), torch.tensor(gts).cuda()
    for att_only in (True, False):
        hp = make_hp(max_N=N + 3, max_T=T + 2, dropout_rate=0.0, attention_guide_dir="from-the-batch", attention_guide_fa=fa)
        if att_only:
            hp.lw_mel, hp.lw_bd1, hp.lw_att, hp.lw_t2m_l2 = 0.0, 0.0, 1.0, 0.0
        P = oracle_params(hp, "t2m", seed=4)
        Pt = ot.to_torch(P, torch.float64, requires_grad=True)
        opt = ot.TFAdam(hp, Pt)
        g = _graph(hp, "train", P, data=iter([]))
        for step in range(1 if att_only else 2):
            comps_ref, grads_ref = ot.text2mel_train_step(hp, Pt, opt, L, mels, gts=gts)
            comps = g.train_step_device(Ld, md, gd).cpu().numpy()
            np.testing.assert_allclose(comps, comps_ref, rtol=2e-4, atol=1e-6)
            if att_only:
                sd = g.store.grads
                check_grads({n: sd[n].cpu().numpy() for n in grads_ref}, grads_ref, "Text2Mel/AudioEnc/HC_13/")
    # a batch without targets is refused when the configuration promises them, and the other way round
    with pytest.raises(AssertionError):
        g.train_step_device(Ld, md)


def test_train_step_with_confidence_losses_matches_oracle():
    """hp.lw_cdp / lw_ain / lw_aout (architectures.py:283-321): coverage deviation penalty and the two attention entropies
    as training losses.  Alone first (their gradient reaches the encoders through Q and K only: tight bound on the
    ReLU-free AudioEnc tail), then with every term over two optimiser steps, then under the loss_weights dict pattern,
    where the reference reports them but leaves them out of the total (:325-331)."""
    B, N, T = 3, 37, 90
    b = synthetic_batch(make_hp(), B, N, T, ragged=True)
    L, mels = torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64)
    Ld, md = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()
    for case in ("alone", "all", "dict"):
        hp = make_hp(max_N=N + 3, max_T=T + 2, dropout_rate=0.0, lw_cdp=0.3, lw_ain=0.4, lw_aout=0.2)
        if case == "alone":
            hp.lw_mel, hp.lw_bd1, hp.lw_att, hp.lw_t2m_l2 = 0.0, 0.0, 0.0, 0.0
        if case == "dict":
            hp.loss_weights = {"t2m": {"L1": 0.3, "binary_divergence": 0.3, "attention": 0.3, "L2": 0.1}}
        P = oracle_params(hp, "t2m", seed=8)
        Pt = ot.to_torch(P, torch.float64, requires_grad=True)
        opt = ot.TFAdam(hp, Pt)
        g = _graph(hp, "train", P, data=iter([]))
        for step in range(1 if case == "alone" else 2):
            comps_ref, grads_ref = ot.text2mel_train_step(hp, Pt, opt, L, mels)
            comps = g.train_step_device(Ld, md).cpu().numpy()
            assert len(comps) == len(comps_ref) == 8
            np.testing.assert_allclose(comps, comps_ref, rtol=2e-4, atol=1e-6)
            if case == "alone":
                sd = g.store.grads
                check_grads({n: sd[n].cpu().numpy() for n in grads_ref}, grads_ref, "Text2Mel/AudioEnc/HC_13/")
        if case == "dict":
            assert abs(comps[0] - (0.3 * comps[1] + 0.3 * comps[2] + 0.3 * comps[3] + 0.1 * comps[4])) < 1e-5


def test_autoregressive_loop_matches_oracle():
    from ophelia_b200.session import Session
    from ophelia_b200 import synthesize as syn
    hp = make_hp(max_N=24, max_T=14)
    P = oracle_params(hp, "t2m", seed=3)
    b = synthetic_batch(hp, 3, 24, 14, text_len=5)
    g = _graph(hp, "synthesize", P)
    sess = Session()
    K, V = syn.encode_text(hp, b["L"], g, sess)
    ends = syn.get_text_lengths(b["L"])
    Kr, Vr = on.TextEnc(hp, P, b["L"])
    Yr, tr, ar = on.synth_codedtext2mel(hp, P, Kr, Vr, ends)
    Y1, t1, a1 = syn.synth_codedtext2mel(hp, K, V, ends, g, sess)
    Y2, t2, a2 = syn.synth_codedtext2mel_device(hp, K, V, ends, g)
    assert t1 == tr and t2 == tr
    assert maxabs(Y1, Yr) < 1e-3 and maxabs(Y2, Yr) < 1e-3
    assert maxabs(a1, ar) < 1e-4 and maxabs(a2, ar) < 1e-4
    Y3, t3 = syn.synth_text2mel(hp, b["L"], g, sess)
    assert t3 == tr and maxabs(Y3, Yr) < 1e-3
    # incremental route: cached AudioEnc rows (fp32 frame-step kernels), Attention / AudioDec over the decoder's reach
    for kw in (dict(use_cuda_graph=True), dict(use_cuda_graph=False, check_every=1), dict(fused_encoder=False)):
        Y5, t5, a5 = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g, **kw)
        assert t5 == tr and maxabs(Y5, Yr) < 1e-3 and maxabs(a5, ar) < 1e-4
    # early stop (synthesize.py:225-228): with the sentence ends at position 1 every sentence ends within a few frames;
    # the device route checks only every 4th frame and must clear the frames it computed past the stopping frame
    ends1 = np.ones_like(ends)
    Yr, tr, ar = on.synth_codedtext2mel(hp, P, Kr, Vr, ends1)
    assert max(tr) < hp.max_T - 1, tr
    for kw in (dict(use_cuda_graph=True, check_every=4), dict(use_cuda_graph=False, check_every=1)):
        Y4, t4, a4 = syn.synth_codedtext2mel_device(hp, K, V, ends1, g, **kw)
        assert t4 == tr and maxabs(Y4, Yr) < 1e-3 and maxabs(a4, ar) < 1e-4
        assert (Y4[:, max(tr) + 1:] == 0).all() and (a4[:, :, max(tr) + 1:] == 0).all()
        Y6, t6, a6 = syn.synth_codedtext2mel_incremental(hp, K, V, ends1, g, **kw)
        assert t6 == tr and maxabs(Y6, Yr) < 1e-3 and maxabs(a6, ar) < 1e-4
        assert (Y6[:, max(tr) + 1:] == 0).all() and (a6[:, :, max(tr) + 1:] == 0).all()


def test_windowless_synthesis_on_the_captured_device_route():
    """hp.turn_off_monotonic_for_synthesis through synth_codedtext2mel_device with a CUDA graph (what the synthesize()
    driver uses for this configuration): the per-sentence key counts of networks.py:307-309 must come from the batch
    being synthesised -- also when a second batch of the same shape replays the captured graph."""
    from ophelia_b200 import synthesize as syn
    from ophelia_b200.session import Session
    hp = make_hp(max_N=24, max_T=12)
    hp.turn_off_monotonic_for_synthesis = True
    P = oracle_params(hp, "t2m", seed=4)
    g = _graph(hp, "synthesize", P)
    sess = Session()
    for text_len in (9, 17):                        # same shapes, different sentence lengths
        b = synthetic_batch(hp, 3, 24, 12, text_len=text_len, seed=text_len)
        hp.text_lengths = syn.get_text_lengths(b["L"]) + 1
        K, V = syn.encode_text(hp, b["L"], g, sess)
        ends = np.array([hp.max_N + 1] * 3)
        Yr, tr, ar = syn.synth_codedtext2mel(hp, K, V, ends, g, sess)          # eager Session route (pinned to the oracle
        Yd, td, ad = syn.synth_codedtext2mel_device(hp, K, V, ends, g, use_cuda_graph=True)   # in test_synthesis_mode_window_mask)
        assert td == tr and maxabs(Yd, Yr) < 1e-5 and maxabs(ad, ar) < 1e-6
        assert (ad[:, text_len + 2:, :] == 0).all() and (ad[:, :text_len + 1, :] > 0).all()
    assert len(g._ar_state) == 1                    # the second batch replayed the first batch's graph


def test_synthesis_at_baseline_shape_with_advancing_attention():
    """BASELINE.json configs[3] (10 sentences, max_N=150, max_T=200, window on) against the fp64 torch oracle loop
    (synthesize.py:150-230), every route.  Random-init weights never move the attention, so the keys carry the synthetic
    diagonal bias of oracle.dctts_torch.advancing_keys (SURVEY 8(d), C4 parity variant): the argmax walks through the
    sentences at different paces, six sentences end in the middle of the run (t_ends between 72 and 194), four run all 200
    frames.  Each frame's argmax feeds the next frame's window, so one wrong decision would derail everything after it:
    the oracle reports its smallest top-2 score gap (1.8e-4 here), an order above the score error of the GPU path."""
    from ophelia_b200 import synthesize as syn
    from ophelia_b200.session import Session
    B, N, T = 10, 150, 200
    hp = make_hp(max_N=N, max_T=T)
    P = oracle_params(hp, "t2m", seed=3)
    rng = np.random.default_rng(1234)
    L = np.zeros((B, N), np.int32)
    for i in range(B):
        n = int(rng.integers(40, N - 1))
        L[i, :n] = rng.integers(1, len(hp.vocab), n)
    ends = syn.get_text_lengths(L)
    Pt = ot.to_torch(P, torch.float64)
    with torch.no_grad():
        Kt, Vt = ot.TextEnc(hp, Pt, L)
    K64 = ot.advancing_keys(Kt.numpy(), c=8.0, seed=0)
    Yr, tr, ar, margin = ot.synth_codedtext2mel(hp, Pt, torch.tensor(K64), Vt, ends, return_margin=True)
    assert margin > 1e-4, margin
    assert sum(t < T for t in tr) >= 4 and sum(t == T for t in tr) >= 2, tr        # early ends and full-length sentences
    K, V = K64.astype(np.float32), Vt.numpy().astype(np.float32)
    g = _graph(hp, "synthesize", P)
    routes = {
        "session": lambda: syn.synth_codedtext2mel(hp, K, V, ends, g, Session()),
        "device": lambda: syn.synth_codedtext2mel_device(hp, K, V, ends, g, use_cuda_graph=False),
        "device_graph": lambda: syn.synth_codedtext2mel_device(hp, K, V, ends, g, use_cuda_graph=True),
        "incremental": lambda: syn.synth_codedtext2mel_incremental(hp, K, V, ends, g),
        "incremental_layers": lambda: syn.synth_codedtext2mel_incremental(hp, K, V, ends, g, fused_encoder=False),
    }
    for name, run in routes.items():
        Y, t, a = run()
        assert t == tr, (name, t, tr)
        assert maxabs(Y, Yr) < 1e-3, (name, maxabs(Y, Yr))
        assert maxabs(a, ar) < 1e-4, (name, maxabs(a, ar))
        assert (a.argmax(1) == ar.argmax(1))[a.max(1) > 0].all(), name                # the whole attention path


def test_incremental_route_beyond_the_decoder_reach():
    """max_T = 120 > 85: the Attention / AudioDec window slides (rows [j - 84, j]) and the widest AudioEnc dilation
    (2 * 27 frames back) reads real history.  The numpy oracle's O(T^2) loop is too slow here, so the reference is the
    full re-computation route on the GPU, itself pinned to the oracle above; the two routes differ only by rounding
    (fp32 FMA vs split-bf16 AudioEnc).  Replaying the cached graph is bit-reproducible."""
    from ophelia_b200 import synthesize as syn
    hp = make_hp(max_N=30, max_T=120)
    P = oracle_params(hp, "t2m", seed=6)
    b = synthetic_batch(hp, 2, 30, 120, text_len=24)
    g = _graph(hp, "synthesize", P)
    enc = g.encode_text({"L": b["L"]})
    K, V = enc["K"].cpu().numpy(), enc["V"].cpu().numpy()
    ends = np.array([hp.max_N + 1] * 2)                              # never reached: all 120 frames are generated
    Yf, tf_, af = syn.synth_codedtext2mel_device(hp, K, V, ends, g)
    Yi, ti, ai = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g)
    assert ti == tf_ == [120, 120] and maxabs(Yi, Yf) < 1e-3 and maxabs(ai, af) < 1e-4
    assert np.abs(Yf[:, 100:]).max() > 0
    assert len(set(af[0].argmax(0).tolist())) > 2                     # the attention window moved during the run
    Y2, t2, a2 = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g)
    assert t2 == ti and np.array_equal(Y2, Yi) and np.array_equal(a2, ai)
    # two launches per AudioEnc layer instead of the one-launch cluster kernel: same function, another summation order
    Y3, t3, a3 = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g, fused_encoder=False)
    assert t3 == ti and maxabs(Y3, Yi) < 1e-4 and maxabs(a3, ai) < 1e-5


def test_side_streams_and_cuda_graph_match_single_stream():
    """TextEnc / weight-gradient side streams and whole-step CUDA-graph replay run the same kernels as the in-order
    eager step: losses and updated weights agree up to the order of the split-K / RED accumulations."""
    import copy
    B, N, T = 3, 60, 200
    hp1 = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    hp2 = copy.copy(hp1)
    hp1.use_side_streams = False
    hp2.use_side_streams = True
    P = oracle_params(hp1, "t2m", seed=5)
    b = synthetic_batch(hp1, B, N, T, ragged=True)
    Ld, md = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()
    g1 = _graph(hp1, "train", P, data=iter([]))
    g2 = _graph(hp2, "train", P, data=iter([]))
    for _ in range(2):
        c1 = g1.train_step_device(Ld, md).cpu().numpy()
        c2 = g2.train_step_device(Ld, md).cpu().numpy()
        np.testing.assert_allclose(c2, c1, rtol=1e-5, atol=1e-7)
    step = g2.capture_train_step(Ld, md, warmup=0)
    c1 = g1.train_step_device(Ld, md).cpu().numpy()
    c2 = step(Ld, md).cpu().numpy()
    np.testing.assert_allclose(c2, c1, rtol=1e-5, atol=1e-7)
    assert int(g1.store.global_step.item()) == int(g2.store.global_step.item()) == 3
    s1, s2 = g1.store.state_dict(), g2.store.state_dict()
    for name in s1:
        assert maxabs(s1[name], s2[name]) < 2e-5, name


def test_capture_survives_dead_cycle_owning_another_captured_step():
    """A discarded Graph object is a reference cycle; the autoregressive route parks a captured step (with its private
    pool) on it.  If Python's cyclic collector frees that cycle while the next capture is open, the release invalidates
    the capture (cudaErrorStreamCaptureInvalidated at the next launch).  ops.capture collects first and keeps the
    collector off while the capture is open; here the collector is off from the start, so the garbage survives until
    ops.capture itself removes it -- before the capture begins."""
    import gc
    import weakref
    from ophelia_b200 import synthesize as syn
    B, N, T = 2, 20, 32
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    P = oracle_params(hp, "t2m", seed=2)
    b = synthetic_batch(hp, B, N, T, ragged=False)
    Ld, md = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()
    gc.collect()
    gc.disable()
    try:
        gs = _graph(hp, "synthesize", P)
        enc = gs.encode_text({"L": b["L"]})
        K, V = enc["K"].cpu().numpy(), enc["V"].cpu().numpy()
        syn.synth_codedtext2mel_device(hp, K, V, [N + 1] * B, gs, use_cuda_graph=True)
        assert gs._ar_state
        dead = weakref.ref(gs)
        del gs, enc                                  # garbage that only the cyclic collector can free (Node <-> Graph)
        g = _graph(make_hp(max_N=N, max_T=T, dropout_rate=0.0), "train", P, data=iter([]))
        g.train_step_device(Ld, md)                  # lazy initialisation outside the capture
        assert dead() is not None
        step = g.capture_train_step(Ld, md, warmup=0)
        assert dead() is None and not gc.isenabled()
        c = step(Ld, md).cpu().numpy()
    finally:
        gc.enable()
    assert np.isfinite(c).all()


def test_full_size_matches_oracle_b32_n180_t870():
    """BASELINE.json configs[1] at its own size (B=32, N=180, T=870) against the fp64 torch oracle: forward outputs of the
    un-masked graph (mels <= 1e-3, attention <= 1e-4, max_attentions exact except the oracle's near-ties), then one
    dropout-free optimiser step (loss components rel 2e-4, gradient of every network in Frobenius norm).  At this size
    the persistent GEMM runs several rounds per launch and reuses both accumulator stages, the remainder K-split and the
    split-K weight gradients see their production schedules (architectures.py:188-362, train.py:273)."""
    from ophelia_b200.session import Session
    B, N, T = 32, 180, 870
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    P = oracle_params(hp, "t2m", seed=7)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Lt, mt = torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64)
    out = ot.text2mel_forward(hp, Pt, Lt, mt, "train")
    comps_t = ot.text2mel_loss(hp, out, mt)
    comps_t[0].backward()
    grads_ref = {k: p.grad.numpy() for k, p in Pt.items() if p.grad is not None}
    comps_ref = [float(c.detach()) for c in comps_t]
    ref = {k: v.detach().numpy() for k, v in out.items()}
    # forward (generate_attention mode = the training graph without dropout, no window mask)
    store = None
    g = _graph(hp, "generate_attention", P)
    Y, ali, mx, K, V = Session().run([g.Y, g.alignments, g.max_attentions, g.K, g.V], {g.L: b["L"], g.mels: b["mels"]})
    assert maxabs(K, ref["K"]) < 1e-3 and maxabs(V, ref["V"]) < 1e-3
    assert maxabs(Y, ref["Y"]) < 1e-3, maxabs(Y, ref["Y"])
    assert maxabs(ali, ref["alignments"]) < 1e-4, maxabs(ali, ref["alignments"])
    bad, ties = argmax_mismatches(mx, ref["max_attentions"], np.swapaxes(ref["alignments"], 1, 2))
    assert bad == 0, (bad, ties)
    del g
    # one optimiser step
    gt = _graph(hp, "train", P, data=iter([]))
    comps = gt.train_step_device(torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(comps, comps_ref, rtol=2e-4, atol=1e-6)
    sd = gt.store.grads
    ours = {n: sd[n].cpu().numpy() for n in grads_ref}
    nets = network_grad_errors(ours, grads_ref, ["Text2Mel/TextEnc/", "Text2Mel/AudioEnc/", "Text2Mel/AudioDec/",
                                                 "Text2Mel/AudioDec/C_11/"])
    # ReLU / |.| kinks flip for a few units at the 2e-5 forward noise level (helpers.check_grads); at this batch size
    # their share is far smaller than at B=2
    assert nets["Text2Mel/AudioDec/C_11/"] < 1e-2, nets
    assert max(nets.values()) < 2e-2, nets
    check_grads(ours, grads_ref, "Text2Mel/AudioDec/C_11/", strict_tol=1e-2)
    opt.step({k: torch.tensor(v) for k, v in grads_ref.items()})
    new = gt.store.state_dict()
    for name, p in Pt.items():
        assert maxabs(new[name], p.detach().numpy()) < 5e-6 + 0.2 * ot.noam(hp.lr, 0), name


def test_full_size_properties_b32_n180_t870():
    """The same shape through properties the reference's graph has by construction: utterances do not interact
    (architectures.py:188-239 has no cross-batch op; here tiles, taps and the boundary fix-up run across item
    boundaries), AudioEnc / AudioDec are causal in time (networks.py:214-284, 360-435), attention rows are
    distributions, and padded text positions embed to zero (modules.py:38-40)."""
    from ophelia_b200.session import Session
    B, N, T = 32, 180, 870
    hp = make_hp(max_N=N, max_T=T)
    P = oracle_params(hp, "t2m", seed=7)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    g = _graph(hp, "generate_attention", P)
    sess = Session()
    Y, ali, Q = sess.run([g.Y, g.alignments, g.Q], {g.L: b["L"], g.mels: b["mels"]})
    assert Y.shape == (B, T, 80) and np.isfinite(Y).all()
    np.testing.assert_allclose(ali.sum(axis=1), 1.0, atol=2e-5)                    # softmax over the N keys
    for i in (0, 17, 31):                                                            # one utterance at a time
        Yi, ai = sess.run([g.Y, g.alignments], {g.L: b["L"][i:i + 1], g.mels: b["mels"][i:i + 1]})
        # another batch size means another tile / split-K schedule, i.e. another summation order of the same fp32-grade
        # arithmetic (differences at its 2e-5 noise level); leakage between utterances would show at the 1e-1 level
        assert maxabs(Yi[0], Y[i]) < 1e-4 and maxabs(ai[0], ali[i]) < 2e-5
    m2 = b["mels"].copy()
    m2[:, 500:] = 1.0 - m2[:, 500:]                                                  # change the future only
    Y2, Q2 = sess.run([g.Y, g.Q], {g.L: b["L"], g.mels: m2})
    # frame t of the input feeds Q[t+1] onwards (one-frame shift, architectures.py:191)
    # (same schedule in both runs and no atomics in the forward pass: bit-identical)
    assert maxabs(Q2[:, :501], Q[:, :501]) == 0.0 and maxabs(Y2[:, :501], Y[:, :501]) == 0.0
    assert maxabs(Y2[:, 501:], Y[:, 501:]) > 1e-3


def test_update_weights_freezes_everything_else():
    """hp.update_weights (architectures.py:113-120, 436-443): only variables whose name matches one of the patterns are
    given to the optimiser.  Two steps against the oracle restricted to the same variables; every other variable (and its
    Adam slots) must stay bit-identical while global_step still advances."""
    from ophelia_b200.architectures import filter_variables_for_update
    B, N, T = 2, 30, 70
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0)
    hp.update_weights = ["Text2Mel/AudioDec", "Text2Mel/TextEnc/HC_1[45]/"]
    P = oracle_params(hp, "t2m", seed=12)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    g = _graph(hp, "train", P, data=iter([]))
    chosen = filter_variables_for_update(g.store, hp.update_weights)
    assert chosen and all(n.startswith("Text2Mel/AudioDec/") or "/TextEnc/HC_14/" in n or "/TextEnc/HC_15/" in n for n in chosen)
    assert len(g.train_ranges) == 2                              # two contiguous runs of the flat buffer
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    Ld, md = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()
    before = g.store.state_dict()
    for _ in range(2):
        for p in Pt.values():
            p.grad = None
        out = ot.text2mel_forward(hp, Pt, torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"], dtype=torch.float64), "train")
        comps_t = ot.text2mel_loss(hp, out, torch.tensor(b["mels"], dtype=torch.float64))
        comps_t[0].backward()
        opt.step({k: Pt[k].grad for k in chosen})               # compute_gradients(loss, var_list=train_variables)
        comps = g.train_step_device(Ld, md).cpu().numpy()
        np.testing.assert_allclose(comps, [float(c.detach()) for c in comps_t], rtol=3e-4, atol=1e-6)
    after = g.store.state_dict()
    assert int(g.store.global_step.item()) == 2
    for name in after:
        if name in chosen:
            assert maxabs(after[name], Pt[name].detach().numpy()) < 5e-6 + 0.2 * 2 * ot.noam(hp.lr, 1), name
            assert maxabs(after[name], before[name]) > 0, name
        else:
            assert np.array_equal(after[name], before[name]), name
            o, cnt = g.store.offsets[name], after[name].size
            assert float(g.store.m_flat[o:o + cnt].abs().max()) == 0.0 and float(g.store.v_flat[o:o + cnt].abs().max()) == 0.0


def test_session_train_loop_with_changing_batch_shapes():
    """train.py:273 through Session.run with the dynamically padded batches of data_load.py:534-541: every batch has its
    own (N_b, T_b).  Shapes seen three times get a CUDA graph, the others run eagerly, batches are prefetched one step
    ahead; the loss trajectory must equal the one of a graph-free, prefetch-free run on the same batches."""
    import copy
    from ophelia_b200.session import Session
    hp1 = make_hp(max_N=40, max_T=120, dropout_rate=0.0)
    hp2 = copy.copy(hp1)
    hp2.use_cuda_graph = False
    hp2.use_side_streams = False
    P = oracle_params(hp1, "t2m", seed=9)
    shapes = [(3, 30, 100), (2, 40, 120), (3, 30, 100), (3, 30, 100), (2, 40, 120), (3, 30, 100), (2, 40, 120),
              (3, 30, 100), (4, 17, 33), (2, 40, 120)]
    batches = []
    for i, (B, N, T) in enumerate(shapes):
        b = synthetic_batch(hp1, B, N, T, seed=100 + i, ragged=True)
        batches.append({"text": torch.tensor(b["L"]), "mel": torch.tensor(b["mels"])})
    g1 = _graph(hp1, "train", P, data=iter(batches))
    g2 = _graph(hp2, "train", P, data=None if False else iter([]))
    sess = Session()
    for i, b in enumerate(batches):
        gs, comps, _ = sess.run([g1.global_step, g1.loss_components, g1.train_op])
        ref = g2.train_step_device(b["text"].cuda(), b["mel"].cuda()).cpu().numpy()
        assert gs == i + 1
        np.testing.assert_allclose(np.asarray(comps), ref, rtol=2e-5, atol=1e-7)
    assert len(g1._graph_steps) == 2                       # the two recurring shapes were captured, (4,17,33) was not
    with pytest.raises(StopIteration):
        sess.run([g1.global_step, g1.loss_components, g1.train_op])


def test_session_returns_at_the_loss_event_and_later_reads_see_the_finished_step():
    """A training Session.run returns when the loss components are on the host (end of the forward pass, an external event
    inside the captured step); the backward pass and the optimiser may still be running.  Everything read afterwards must
    see the finished step: the parameters fetched right after each call have moved by the optimiser step (constant learning
    rate 1e-3 here, so a stale read would be off by ~1e-3 per element) and agree with a twin that waits for the whole
    step (hp.early_loss_return = False) up to the summation order of the gradient atomics; losses and global_step agree."""
    import copy
    from ophelia_b200.session import Session
    hp1 = make_hp(max_N=40, max_T=120, dropout_rate=0.0, decay_lr=False)
    hp2 = copy.copy(hp1)
    hp2.early_loss_return = False
    P = oracle_params(hp1, "t2m", seed=5)
    b = synthetic_batch(hp1, 3, 40, 120, seed=3, ragged=True)
    batches = [{"text": torch.tensor(b["L"]), "mel": torch.tensor(b["mels"])} for _ in range(6)]
    g1 = _graph(hp1, "train", P, data=iter(batches))
    g2 = _graph(hp2, "train", P, data=iter(batches))
    sess = Session()
    prev = g1.store.flat.detach().cpu().numpy().copy()
    for i in range(6):                                    # steps 4-6 replay the captured graph
        out1 = sess.run([g1.global_step, g1.loss_components, g1.train_op])
        w1 = g1.store.flat.detach().cpu().numpy().copy()  # stream-ordered behind the rest of the step
        out2 = sess.run([g2.global_step, g2.loss_components, g2.train_op])
        w2 = g2.store.flat.detach().cpu().numpy().copy()
        assert out1[0] == out2[0] == i + 1
        # (step 1 sees identical weights; afterwards the twins drift apart through the order of the gradient atomics and a
        #  learning rate of 1e-3 on sign-like first Adam steps)
        np.testing.assert_allclose(out1[1], out2[1], rtol=1e-6 if i == 0 else 5e-2, atol=1e-6)
        moved = np.abs(w1 - prev)
        assert float(np.mean(moved > 1e-4)) > 0.5, (i, float(np.mean(moved > 1e-4)))      # the step's update is in what we read (a stale read: 0)
        assert float(np.mean(np.abs(w1 - w2) > 1e-4)) < 0.02, (i, float(np.mean(np.abs(w1 - w2) > 1e-4)))
        prev = w1
    assert g1.__dict__.get("_loss_out") is not None and g2.__dict__.get("_loss_out") is None
    assert len(g1._graph_steps) == 1 and len(g2._graph_steps) == 1


def test_norm_none_configuration_matches_oracle():
    """hp.norm = None (the reference's config/project/*.cfg): conv1d / hc run without layer norm (modules.py:47-75 returns
    its input), no normalize variables exist; forward and two optimiser steps against the oracles."""
    from ophelia_b200.session import Session
    B, N, T = 2, 40, 150
    hp = make_hp(max_N=N, max_T=T, dropout_rate=0.0, norm=None)
    P = oracle_params(hp, "t2m", seed=11)
    b = synthetic_batch(hp, B, N, T, ragged=True)
    ref = on.text2mel_forward(hp, P, b["L"], b["mels"], "generate_attention")
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.variables import VariableStore
    store = VariableStore("cuda:0")
    g = Text2MelGraph(hp, mode="generate_attention", store=store)
    assert not any("normalize" in n or "/H1/" in n for n in store.specs)
    store.load_state_dict(P, strict=False)
    Y, ali = Session().run([g.Y, g.alignments], {g.L: b["L"], g.mels: b["mels"]})
    assert maxabs(Y, ref["Y"]) < 1e-3 and maxabs(ali, ref["alignments"]) < 1e-4
    Pt = ot.to_torch(P, torch.float64, requires_grad=True)
    opt = ot.TFAdam(hp, Pt)
    st2 = VariableStore("cuda:0")
    gt = Text2MelGraph(hp, mode="train", store=st2, data=iter([]))
    st2.load_state_dict(P, strict=False)
    Ld, md = torch.tensor(b["L"]).cuda(), torch.tensor(b["mels"]).cuda()
    for _ in range(2):
        comps_ref, _ = ot.text2mel_train_step(hp, Pt, opt, torch.tensor(b["L"].astype(np.int64)),
                                              torch.tensor(b["mels"], dtype=torch.float64))
        comps = gt.train_step_device(Ld, md).cpu().numpy()
        np.testing.assert_allclose(comps, comps_ref, rtol=5e-4, atol=1e-6)
