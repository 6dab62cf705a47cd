"""Synthesis drivers with the reference's names and semantics (`synthesize.py:62-260`):
synth_text2mel, synth_codedtext2mel, encode_text, get_text_lengths, synth_mel2mag (+ split_batch).

Two routes produce identical results:
  * the Session route keeps the reference's call pattern (one `sess.run` per mel frame, numpy in/out);
  * `synth_codedtext2mel_device` keeps Y / alignments / prev_max_attentions resident in HBM and only reads the
    B argmax values back per frame for the reference's end-of-sentence test.
Both re-run AudioEnc + Attention + AudioDec over all max_T frames every step, because the reference applies the
monotonic window derived from the *latest* prev_max_attentions to every time row (networks.py:304-313).
`synthesize()` / `main_work()` keep the reference's command line (`python -m ophelia_b200.synthesize -c CONFIG [-N n]
[-t2m_epoch e] [-ssrn_epoch e] [-odir dir] ...`): restore both models, Text2Mel loop, SSRN, attention diagnostics,
Griffin-Lim on the GPU (ophelia_b200/vocoder.py).  The WORLD vocoder and babbling are outside the path.
"""
import numpy as np
import torch


def get_text_lengths(L):
    """Index of the first padding symbol (0) in each row (synthesize.py:242-247)."""
    return np.array([int(np.where(L[i, :] == 0)[0][0]) for i in range(len(L))])


def _variant_feeds(hp, g, feeddict, speaker_data=None, duration_data=None, labels=None):
    """The optional placeholders of synthesize.py:83-92 / 173-178."""
    if hp.multispeaker:
        feeddict[g.speakers] = speaker_data
    if hp.use_external_durations and duration_data is not None:
        feeddict[g.durations] = duration_data
    if hp.merlin_label_dir and labels is not None:
        feeddict[g.merlin_label] = labels
    return feeddict


def encode_text(hp, L, g, sess, speaker_data=None, labels=None):
    """One text encoder pass -> (K, V) (synthesize.py:232-240)."""
    K, V = sess.run([g.K, g.V], _variant_feeds(hp, g, {g.L: L}, speaker_data, None, labels))
    return (K, V)


def _update_ends(hp, max_att_j, ends, endcounts, t_ends, j, endcount_threshold=1):
    """End-of-sentence bookkeeping of synthesize.py:218-228: a sentence ends at the first frame whose attention
    argmax reaches the first padding position; returns True when every sentence has ended."""
    endcounts += (max_att_j >= ends)
    for i in range(len(t_ends)):
        if t_ends[i] == hp.max_T and endcounts[i] >= endcount_threshold:
            t_ends[i] = j
    return bool((t_ends < hp.max_T).all())


def synth_text2mel(hp, L, g, sess, speaker_data=None, duration_data=None, labels=None, position_in_phone_data=None):
    """Route that re-runs TextEnc every frame (synthesize.py:62-132).  Returns (Y, t_ends)."""
    B = len(L)
    Y = np.zeros((B, hp.max_T, hp.n_mels), np.float32)
    prev_max_attentions = np.zeros((B,), np.int32)
    ends = get_text_lengths(L)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * hp.max_T
    feeddict = _variant_feeds(hp, g, {g.L: L, g.mels: Y, g.prev_max_attentions: prev_max_attentions}, speaker_data,
                              duration_data, labels)
    for j in range(hp.max_T):
        _Y, _max_attentions, _alignments = sess.run([g.Y, g.max_attentions, g.alignments], feeddict)
        Y[:, j, :] = _Y[:, j, :]
        prev_max_attentions = _max_attentions[:, j]
        feeddict[g.mels] = Y
        feeddict[g.prev_max_attentions] = prev_max_attentions
        if _update_ends(hp, _max_attentions[:, j], ends, endcounts, t_ends, j):
            break
    return (Y, t_ends.tolist())


def synth_codedtext2mel(hp, K, V, ends, g, sess, speaker_data=None, duration_data=None, labels=None,
                        position_in_phone_data=None):
    """Route with K, V encoded once and fed (synthesize.py:150-230).  Returns (Y, t_ends, alignments).  With external
    durations the sentence ends are known in advance (`t_ends = duration_data.sum(axis=(1,2))`, :168-169)."""
    B = len(K)
    Y = np.zeros((B, hp.max_T, hp.n_mels), np.float32)
    alignments = np.zeros((len(ends), hp.max_N, hp.max_T), np.float32)
    prev_max_attentions = np.zeros((B,), np.int32)
    ends = np.asarray(ends)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * hp.max_T
    if hp.use_external_durations:
        t_ends = np.asarray(duration_data).sum(axis=(1, 2)).astype(int)
    feeddict = _variant_feeds(hp, g, {g.K: K, g.V: V, g.mels: Y, g.prev_max_attentions: prev_max_attentions}, speaker_data,
                              duration_data, labels)
    for j in range(hp.max_T):
        _Y, _max_attentions, _alignments = sess.run([g.Y, g.max_attentions, g.alignments], feeddict)
        Y[:, j, :] = _Y[:, j, :]
        alignments[:, :, j] = _alignments[:, :, j]
        prev_max_attentions = _max_attentions[:, j]
        feeddict[g.mels] = Y
        feeddict[g.prev_max_attentions] = prev_max_attentions
        if hp.use_external_durations:                        # synthesize.py:211-215
            if j >= t_ends.max():
                break
        elif _update_ends(hp, _max_attentions[:, j], ends, endcounts, t_ends, j):
            break
    return (Y, t_ends.tolist(), alignments)


def synth_babble(hp, g, sess, seed=False, nsamples=16):
    """synthesize.py:134-148: let a BabblerGraph continue from silence, frame by frame.  Returns Y [nsamples, max_T, n_mels]."""
    assert not seed, 'TODO: implement seeding babbler'
    Y = np.zeros((nsamples, hp.max_T, hp.n_mels), np.float32)
    for j in range(hp.max_T):
        _Y, = sess.run([g.Y], {g.mels: Y})
        Y[:, j, :] = _Y[:, j, :]
    return Y


def babble(hp, num_sentences=0, vocode=True):
    """synthesize.py:333-364: babbler + SSRN (+ Griffin-Lim) from the latest checkpoints; returns the output directory."""
    import os
    from . import vocoder
    from .architectures import BabblerGraph, SSRNGraph
    from .session import Session
    if num_sentences == 0:
        num_sentences = 4
    g1 = BabblerGraph(hp, mode="synthesize"); print("Babbler graph loaded")
    g2 = SSRNGraph(hp, mode="synthesize"); print("SSRN graph loaded")
    with Session() as sess:
        babbler_epoch = restore_latest_model_parameters(sess, hp, 'babbler', graph=g1)
        ssrn_epoch = restore_latest_model_parameters(sess, hp, 'ssrn', graph=g2)
        Y = synth_babble(hp, g1, sess, seed=False, nsamples=num_sentences)
        Z = synth_mel2mag(hp, Y, g2, sess)
        if np.isnan(Z).any():
            Z = np.nan_to_num(Z)
        outdir = os.path.join(hp.voicedir, 'synth_babble', '%s_%s' % (babbler_epoch, ssrn_epoch))
        os.makedirs(outdir, exist_ok=True)
        for i, mag in enumerate(Z):
            if vocode:
                wav = vocoder.spectrogram2wav(hp, mag)
                vocoder.write_wav(outdir + "/{:03d}.wav".format(i), wav, hp.sr)
            else:
                np.save(outdir + "/{:03d}.npy".format(i), mag)
    return outdir


def synth_codedtext2mel_device(hp, K, V, ends, g, use_cuda_graph=True, check_every=8):
    """Same loop with every tensor resident on the GPU.  With use_cuda_graph the per-frame forward (AudioEnc + Attention +
    AudioDec over all max_T frames, ~55 kernels) is captured once and replayed: Y and prev_max_attentions are static
    buffers that the loop updates in place.  The attention argmax of every frame is kept on the device and read back
    only every `check_every` frames for the reference's end-of-sentence test (synthesize.py:218-228); frames computed
    past the stopping frame are cleared again, so the results equal the frame-by-frame loop's."""
    dev = g.device
    from . import ops
    K = g._to_device(K, torch.float32)
    V = g._to_device(V, torch.float32)
    B = K.shape[0]
    # static buffers (and the captured step) are kept on the graph object per batch shape: synthesising many batches
    # re-uses them instead of capturing again
    cache = g.__dict__.setdefault("_ar_state", {})
    key = (B, K.shape[1], K.shape[2], hp.max_T, bool(use_cuda_graph), bool(getattr(hp, "turn_off_monotonic_for_synthesis", False)))
    st = cache.get(key)
    if st is not None and st["version"] != g.store.version:      # weights changed since the capture: packed images are stale
        st = None
    if st is None:
        st = {"K": torch.empty(K.shape, device=dev, dtype=torch.float32), "V": torch.empty(V.shape, device=dev, dtype=torch.float32),
              "Y": torch.zeros(B, hp.max_T, hp.n_mels, device=dev, dtype=torch.float32),
              "prev": torch.zeros(B, device=dev, dtype=torch.int32), "graph": None, "out": None,
              "version": g.store.version}
        for n in ("K", "V"):                            # planes live next to the static buffers and are refreshed per call
            st[n]._oph_planes = ops.split_planes(st[n])
        cache[key] = st
    Kb, Vb, Y, prev = st["K"], st["V"], st["Y"], st["prev"]
    Kb.copy_(K)
    Vb.copy_(V)
    ops.split_planes(Kb, into=Kb._oph_planes)            # K, V stay constant over the loop: split them once per call
    ops.split_planes(Vb, into=Vb._oph_planes)
    Y.zero_()
    prev.zero_()
    alignments = torch.zeros(B, hp.max_N, hp.max_T, device=dev, dtype=torch.float32)
    history = torch.zeros(hp.max_T, B, device=dev, dtype=torch.int32)      # argmax of frame j at frame j
    ends = np.asarray(ends)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * hp.max_T

    # windowless synthesis (networks.py:307-309): the per-sentence key counts live in a static device buffer that is
    # refreshed here, outside any capture, so a replayed graph masks with THIS batch's lengths
    tl = None
    if getattr(hp, "turn_off_monotonic_for_synthesis", False):
        assert len(hp.text_lengths) == B, "hp.text_lengths must describe the batch (synthesize.py:505-507)"
        if "tl" not in st:
            st["tl"] = torch.zeros(B, device=dev, dtype=torch.int32)
        tl = st["tl"]
        tl.copy_(torch.as_tensor(np.asarray(hp.text_lengths), dtype=torch.int32))

    def forward():
        return g.build_model(None, Y, False, K=Kb, V=Vb, prev_max_attentions=prev, want_alignments=True, text_lengths=tl)
    graph, out = st["graph"], st["out"]
    if use_cuda_graph and graph is None:
        forward()                                       # warm-up outside the capture (lazy weight packing)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with ops.capture(graph):
            out = forward()
        st["graph"], st["out"] = graph, out
    checked = 0
    stop_at = None
    for j in range(hp.max_T):
        if graph is not None:
            graph.replay()
        else:
            out = forward()
        Y[:, j, :].copy_(out["Y"][:, j, :])
        alignments[:, :, j].copy_(out["alignments"][:, :, j])
        prev.copy_(out["max_attentions"][:, j])
        history[j].copy_(prev)
        if (j + 1) % max(1, check_every) == 0 or j == hp.max_T - 1:
            host = history[checked:j + 1].cpu().numpy().astype(np.int64)
            for jj in range(checked, j + 1):
                if _update_ends(hp, host[jj - checked], ends, endcounts, t_ends, jj):
                    stop_at = jj
                    break
            checked = j + 1
            if stop_at is not None:
                break
    if stop_at is not None and stop_at + 1 < hp.max_T:   # the reference never computed these frames
        Y[:, stop_at + 1:, :].zero_()
        alignments[:, :, stop_at + 1:].zero_()
    return (Y.cpu().numpy(), t_ends.tolist(), alignments.cpu().numpy())


def _frame_step_layers(hp):
    """(scope, kind, k, rate, act) of the causal stacks in creation order (networks.py:214-284, 360-435)."""
    enc = [("C_1", "conv", 1, 1, 1), ("C_2", "conv", 1, 1, 1), ("C_3", "conv", 1, 1, 0)]
    enc += [("HC_%d" % (4 + i), "hc", 3, 3 ** (i % 4), 0) for i in range(8)]
    enc += [("HC_12", "hc", 3, 3, 0), ("HC_13", "hc", 3, 3, 0)]
    dec = [("C_1", "conv", 1, 1, 0)]
    dec += [("HC_%d" % (2 + i), "hc", 3, 3 ** i, 0) for i in range(4)]
    dec += [("HC_6", "hc", 3, 1, 0), ("HC_7", "hc", 3, 1, 0)]
    dec += [("C_8", "conv", 1, 1, 1), ("C_9", "conv", 1, 1, 1), ("C_10", "conv", 1, 1, 1), ("C_11", "conv", 1, 1, 0)]
    return enc, dec


def decoder_reach(hp):
    """Causal reach of AudioDec in frames: sum over its layers of (k - 1) * rate = 84 (networks.py:360-435)."""
    return sum((k - 1) * rate for _scope, _kind, k, rate, _act in _frame_step_layers(hp)[1])


def _window_hp(hp, W):
    """networks.Attention asserts T == hp.max_T in synthesis mode (the reference tiles its mask over hp.max_T rows):
    the windowed pass over W rows gets a copy of hp that says so."""
    import copy
    hp_w = copy.copy(hp)
    hp_w.max_T = W
    return hp_w


def synth_codedtext2mel_incremental(hp, K, V, ends, g, use_cuda_graph=True, check_every=8, fused_encoder=None):
    """The same loop with the per-frame work cut to what can change (csrc/arstep.cuh; SURVEY 7, hard part 5):

    * AudioEnc is causal and its input row t is final once frame t - 1 exists: every layer keeps its output history and
      a frame step computes row j only (fp32 FMA kernels that stream the fp32 weights once per step);
    * Attention cannot be cached -- the window mask of the latest prev_max_attentions is applied to every time row
      (networks.py:304-313), so the context rows R[t < j] change whenever the window moves, and AudioDec's row j sees them
      through its 84-frame causal reach.  Attention and AudioDec therefore run per step with the batch (tcgen05) kernels,
      but over the rows [max(0, j - 84), j] only.

    fused_encoder (default: whenever d is 256 or 512): the 13 AudioEnc layers of a frame step run in ONE launch, an
    8-CTA thread-block cluster per sentence (oph_ar_encoder_step), instead of two launches per layer: 0.388 against
    0.413 ms per frame step on B200 (10 sentences), same results up to the summation order.

    One captured CUDA graph is replayed per frame (the frame index lives on the device).  Results equal the full
    re-computation up to fp32 rounding (tests/test_incremental_algorithm.py pins the algorithm against the oracle loop);
    K, V, ends and the return value are those of `synth_codedtext2mel`."""
    from . import _lib, ops
    from .networks import Attention, AudioDec
    from .variables import use_store, variable_scope
    assert not hp.multispeaker and not hp.use_external_durations and not hp.merlin_label_dir
    assert not getattr(hp, "turn_off_monotonic_for_synthesis", False), "windowless attention: use synth_codedtext2mel_device"
    dev, st = g.device, g.store
    K = g._to_device(K, torch.float32).contiguous()
    V = g._to_device(V, torch.float32).contiguous()
    B, N, d = K.shape
    T, nm = hp.max_T, hp.n_mels
    assert N == hp.max_N, "networks.py:304-311 builds the mask with hp.max_N"
    assert B <= 16, "the frame-step kernels handle up to 16 sentences per call"
    assert hp.norm in ('layer', None)
    norm = hp.norm == 'layer'
    reach = decoder_reach(hp)
    W = min(T, reach + 1)
    hp_w = _window_hp(hp, W)
    enc_layers = _frame_step_layers(hp)[0]
    if fused_encoder is None:
        fused_encoder = d in (256, 512)
    assert not fused_encoder or d in (256, 512), "oph_ar_encoder_step needs d = 256 or 512"
    cache = g.__dict__.setdefault("_ar_inc_state", {})
    key = (B, N, d, T, bool(use_cuda_graph), bool(fused_encoder), st.flat.data_ptr())
    state = cache.get(key)
    if state is not None and state["version"] != st.version:    # the windowed decoder reads packed weight images
        state = None
    if state is None:
        z = lambda *shape, dt=torch.float32: torch.zeros(*shape, device=dev, dtype=dt)
        state = {"K": z(B, N, d), "V": z(B, N, d), "Y": z(B, T, nm), "ali": z(B, N, T), "prev": z(B, dt=torch.int32),
                 "hist": z(T, B, dt=torch.int32), "frame": z(1, dt=torch.int32), "Qw": z(B, W, d),
                 "scratch": z(int(_lib.load().oph_ar_scratch_floats())), "enc": [z(B, T, d) for _ in enc_layers],
                 "graph": None, "version": st.version}
        for n in ("K", "V"):
            state[n]._oph_planes = ops.split_planes(state[n])
        cache[key] = state
    Kb, Vb, Y, ali, prev, hist, frame, Qw = (state[n] for n in ("K", "V", "Y", "ali", "prev", "hist", "frame", "Qw"))

    def var(name):
        return st.vars.get("Text2Mel/AudioEnc/" + name)

    def enc_layer(spec, x, y, in_shift):
        scope, kind, k, rate, act = spec
        w, bias = var(scope + "/conv1d/kernel"), var(scope + "/conv1d/bias")
        stream = torch.cuda.current_stream().cuda_stream
        if kind == "hc":
            C = x.shape[2]
            assert tuple(w.shape) == (k, C, 2 * C) and y.shape[2] == C
            g1, b1, g2, b2 = (var(scope + n) if norm else None for n in ("/H1/gamma", "/H1/beta", "/H2/gamma", "/H2/beta"))
            _lib.call("oph_ar_hc_step", x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), bias.data_ptr(),
                      ops._p(g1), ops._p(b1), ops._p(g2), ops._p(b2), y.data_ptr(), y.stride(0), y.stride(1),
                      state["scratch"].data_ptr(), B, C, k, rate, frame.data_ptr(), stream)
        else:
            cin, cout = x.shape[2], y.shape[2]
            assert tuple(w.shape) == (k, cin, cout)
            gamma, beta = (var(scope + n) if norm else None for n in ("/normalize/gamma", "/normalize/beta"))
            _lib.call("oph_ar_conv_step", x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), bias.data_ptr(),
                      ops._p(gamma), ops._p(beta), y.data_ptr(), y.stride(0), y.stride(1), None, 0, 0,
                      state["scratch"].data_ptr(), B, cin, cout, k, rate, in_shift, act, frame.data_ptr(), stream)

    def encoder_table():
        """oph_ar_layer[13]: the static description of the AudioEnc frame step for the one-launch kernel."""
        table = (_lib.ArLayer * len(enc_layers))()
        x = Y
        for i, (scope, kind, k, rate, act) in enumerate(enc_layers):
            y = state["enc"][i]
            e = table[i]
            names = ("/H1/gamma", "/H1/beta", "/H2/gamma", "/H2/beta") if kind == "hc" else ("/normalize/gamma", "/normalize/beta")
            params = [ops._p(var(scope + n)) if norm else None for n in names] + [None, None]
            e.w, e.bias = var(scope + "/conv1d/kernel").data_ptr(), var(scope + "/conv1d/bias").data_ptr()
            e.g1, e.b1, e.g2, e.b2 = params[:4]
            e.x, e.y = x.data_ptr(), y.data_ptr()
            e.x_item, e.ldx, e.y_item, e.ldy = x.stride(0), x.stride(1), y.stride(0), y.stride(1)
            e.Cin, e.C, e.k, e.rate = x.shape[2], y.shape[2], k, rate
            e.kind, e.act, e.in_shift = int(kind == "hc"), act, 1 if i == 0 else 0
            x = y
        return table

    def step():
        x = Y                                              # S = mels delayed by one frame (architectures.py:191)
        if fused_encoder:
            if "enc_table" not in state:
                state["enc_table"] = encoder_table()
            _lib.call("oph_ar_encoder_step", state["enc_table"], len(enc_layers), B, frame.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
            x = state["enc"][-1]
        else:
            for i, spec in enumerate(enc_layers):
                enc_layer(spec, x, state["enc"][i], 1 if i == 0 else 0)
                x = state["enc"][i]
        stream = torch.cuda.current_stream().cuda_stream
        _lib.call("oph_ar_window_gather", x.data_ptr(), x.stride(0), x.stride(1), Qw.data_ptr(), Qw.stride(0), Qw.stride(1),
                  B, d, T, W, reach, frame.data_ptr(), stream)
        with use_store(st), variable_scope("Text2Mel"):
            with variable_scope("Attention"):
                Rw, ali_w, max_w = Attention(hp_w, Qw, Kb, Vb, monotonic_attention=True, prev_max_attentions=prev,
                                             training=False, want_alignments=True)
            with variable_scope("AudioDec"):
                _logits_w, Yw = AudioDec(hp, Rw, training=False, speaker_codes=None, reuse=g.reuse)
        assert ali_w.is_contiguous() and max_w.is_contiguous() and max_w.dtype == torch.int32
        _lib.call("oph_ar_window_scatter", Yw.data_ptr(), Yw.stride(0), Yw.stride(1), Y.data_ptr(), Y.stride(0), Y.stride(1),
                  nm, ali_w.data_ptr(), ali.data_ptr(), max_w.data_ptr(), prev.data_ptr(), hist.data_ptr(), B, N, T, W,
                  reach, frame.data_ptr(), stream)
        _lib.call("oph_ar_advance", frame.data_ptr(), stream)

    def reset():
        Kb.copy_(K)
        Vb.copy_(V)
        ops.split_planes(Kb, into=Kb._oph_planes)
        ops.split_planes(Vb, into=Vb._oph_planes)
        for t in (Y, ali, prev, hist, frame, state["enc"][-1]):
            t.zero_()
    graph = state["graph"]
    if use_cuda_graph and graph is None:
        reset()
        step()                                             # warm-up outside the capture (lazy weight packing)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with ops.capture(graph):
            step()
        state["graph"] = graph
    reset()
    ends = np.asarray(ends)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * T
    checked, stop_at = 0, None
    for j in range(T):
        if graph is not None:
            graph.replay()
        else:
            step()
        if (j + 1) % max(1, check_every) == 0 or j == T - 1:
            host = hist[checked:j + 1].cpu().numpy().astype(np.int64)
            for jj in range(checked, j + 1):
                if _update_ends(hp, host[jj - checked], ends, endcounts, t_ends, jj):
                    stop_at = jj
                    break
            checked = j + 1
            if stop_at is not None:
                break
    Yh, ah = Y.cpu().numpy(), ali.cpu().numpy()
    if stop_at is not None and stop_at + 1 < T:           # the reference never computed these frames
        Yh[:, stop_at + 1:, :] = 0
        ah[:, :, stop_at + 1:] = 0
    return (Yh, t_ends.tolist(), ah)


def synth_codedtext2mel_fast(hp, K, V, ends, g):
    """The fastest device route that applies: incremental for up to 16 sentences with the attention window on, else the
    CUDA-graph replay of the full re-computation.  Same results (up to fp32 rounding) and return value as
    `synth_codedtext2mel`."""
    if len(K) <= 16 and not getattr(hp, "turn_off_monotonic_for_synthesis", False):
        return synth_codedtext2mel_incremental(hp, K, V, ends, g)
    return synth_codedtext2mel_device(hp, K, V, ends, g)


def synth_mel2mag(hp, Y, g, sess, batchsize=128):
    """SSRN over the padded mel batch in chunks of <= batchsize utterances (synthesize.py:250-260)."""
    if batchsize > 0:
        nbatches = max(1, len(Y) // batchsize)
        batches = np.array_split(Y, nbatches)
    else:
        batches = [Y]
    Z = np.concatenate([sess.run(g.Z, {g.mels: Y_batch}) for Y_batch in batches])
    return Z


_MODEL_SCOPES = {'t2m': 'Text2Mel', 'ssrn': 'SSRN', 'babbler': 'Text2Mel'}     # synthesize.py:303-306


def restore_latest_model_parameters(sess, hp, model_type, graph=None):
    """synthesize.py:302-316: restore the trainable variables of `model_type` from the latest TF checkpoint under
    `hp.logdir + "-" + model_type` (read without TensorFlow, ophelia_b200/tf_checkpoint.py).  `graph` is the
    Text2MelGraph / SSRNGraph whose variable store receives the values (the reference finds its variables through the
    default TF graph).  Returns the epoch string of the checkpoint name."""
    import sys
    from . import tf_checkpoint
    assert model_type in _MODEL_SCOPES, "model type %r is outside the path" % (model_type,)
    savepath = hp.logdir + "-" + model_type
    latest_checkpoint = tf_checkpoint.latest_checkpoint(savepath)
    if latest_checkpoint is None:
        sys.exit('No %s at %s?' % (model_type, savepath))
    latest_epoch = latest_checkpoint.strip('/ ').split('/')[-1].replace('model_epoch_', '')
    tf_checkpoint.restore(graph.store, latest_checkpoint, strict=True, with_optimizer=False)
    print("Model of type %s restored from latest epoch %s" % (model_type, latest_epoch))
    return latest_epoch


def restore_archived_model_parameters(sess, hp, model_type, epoch_number, graph=None):
    """synthesize.py:319-330: same from `<logdir>-<model_type>/archive/model_epoch_<n>`."""
    import os
    import sys
    from . import tf_checkpoint
    assert model_type in _MODEL_SCOPES, "model type %r is outside the path" % (model_type,)
    desired_checkpoint = hp.logdir + "-" + model_type + "/archive/model_epoch_" + str(epoch_number)
    if not os.path.isfile(desired_checkpoint + '.index'):
        sys.exit('No %s at %s?' % (model_type, desired_checkpoint))
    tf_checkpoint.restore(graph.store, desired_checkpoint, strict=True, with_optimizer=False)
    print("Model of type %s restored from archived epoch %s" % (model_type, epoch_number))


def make_mel_batch(hp, fnames, oracle=True):
    """synthesize.py:270-284: natural coarse mels of `fnames` (from hp.coarse_audio_dir when `oracle`, else the paths
    themselves) zero-padded to [n, max_T, n_mels]; lengths in full-rate frames (r per coarse frame)."""
    import os
    import re
    if oracle:
        paths = [os.path.join(hp.coarse_audio_dir, re.sub(r'\.[^\.]+\Z', '', os.path.split(f)[1]) + '.npy') for f in fnames]
    else:
        paths = list(fnames)
    batch = np.zeros((len(paths), hp.max_T, hp.n_mels), np.float32)
    lengths = []
    for i, path in enumerate(paths):
        mel = np.load(path)
        batch[i, :mel.shape[0], :] = mel
        lengths.append(mel.shape[0] * hp.r)
    return batch, lengths


def list2batch(inlist, pad_length):
    """synthesize.py:286-299: stack [len_i, dim] arrays into [n, pad_length (0 = longest), dim], zero padded."""
    dim = inlist[0].shape[1]
    if pad_length == 0:
        pad_length = max(a.shape[0] for a in inlist)
    batch = np.zeros((len(inlist), pad_length, dim), np.float32)
    for i, array in enumerate(inlist):
        assert array.shape[0] <= pad_length and array.shape[1] == dim
        batch[i, :array.shape[0], :] = array
    return batch


def split_batch(synth_batch, end_indices):
    return [predmel[:end_indices[i], :] for i, predmel in enumerate(synth_batch)]


# ---------------------------------------------------------------------------------------------- attention diagnostics
def getCDP(A):
    """calculate_CDP_Ain_Aout.py:18-25: coverage deviation penalty of a trimmed alignment [N, T] (0 = every input symbol
    receives a total attention of exactly 1; trailing symbols without any attention are not counted)."""
    att_per_input = np.trim_zeros(np.sum(A, axis=1), 'b')
    return float(np.sum(np.log(1. + (1. - att_per_input) ** 2)) / len(att_per_input))


def _mean_entropy(A):
    """calculate_CDP_Ain_Aout.py:27-43: mean entropy of the rows of A, each renormalised to sum 1, in units of
    log(row length)."""
    A = np.asarray(A, np.float64)
    norm = A.sum(axis=1, keepdims=True)
    P = np.divide(A, norm, out=A.copy(), where=norm != 0.0)
    plogp = np.where(P != 0.0, P * np.log(np.where(P != 0.0, P, 1.0)), 0.0)
    return float(-plogp.sum() / A.shape[0] / np.log(A.shape[1]))


def getAP(A):
    """calculate_CDP_Ain_Aout.py:46-57: absent-mindedness penalties (APin, APout): attention dispersion per input symbol
    and per output frame."""
    num_input = len(np.trim_zeros(np.sum(A, axis=1), 'b'))
    A = A[:num_input, :]
    return _mean_entropy(A), _mean_entropy(np.transpose(A))


# ---------------------------------------------------------------------------------------------- driver
def synthesize(hp, speaker_id='', num_sentences=0, ncores=1, topoutdir='', t2m_epoch=-1, ssrn_epoch=-1, vocode=True):
    """synthesize.py:442-632 for the attention-driven single-speaker configurations.  Writes `<base>.wav` (and
    `<base>_alignment.npy` in place of the attention plot) under `<topoutdir or hp.sampledir>/t2m<e>_ssrn<e>/`.
    Returns (outdir, bases, lengths).  `ncores` is accepted for command-line compatibility: Griffin-Lim runs on the GPU.
    vocode=False stops after SSRN (BASELINE config 4: "Griffin-Lim off") and stores the magnitudes as .npy."""
    import os
    import re
    import time
    from . import vocoder
    from .architectures import SSRNGraph, Text2MelGraph
    from .data_load import load_data
    from .session import Session
    assert hp.vocoder in ['griffin_lim'], 'Other vocoders than griffin_lim are outside the path'
    assert not speaker_id and not hp.multispeaker, "multi-speaker synthesis is outside the path"
    dataset = load_data(hp, mode="synthesis")
    fpaths, L = dataset['fpaths'], dataset['texts']
    if num_sentences > 0:
        assert num_sentences <= len(fpaths)
        L = L[:num_sentences, :]
        fpaths = fpaths[:num_sentences]
    bases = [re.sub(r'\.[^\.]+\Z', '', os.path.split(f)[1]) for f in fpaths]
    if hp.turn_off_monotonic_for_synthesis:
        hp.text_lengths = get_text_lengths(L) + 1
    g1 = Text2MelGraph(hp, mode="synthesize"); print("Graph 1 (t2m) loaded")
    # synthesize.py:513-533: a Text2Mel trained without layer norm is paired with an SSRN that has it
    hp2 = hp
    if hp.norm is None:
        import copy
        hp2 = copy.copy(hp)
        hp2.norm = 'layer'
    g2 = SSRNGraph(hp2, mode="synthesize"); print("Graph 2 (ssrn) loaded")
    with Session() as sess:
        if t2m_epoch > -1:
            restore_archived_model_parameters(sess, hp, 't2m', t2m_epoch, graph=g1)
        else:
            t2m_epoch = restore_latest_model_parameters(sess, hp, 't2m', graph=g1)
        if ssrn_epoch > -1:
            restore_archived_model_parameters(sess, hp, 'ssrn', ssrn_epoch, graph=g2)
        else:
            ssrn_epoch = restore_latest_model_parameters(sess, hp, 'ssrn', graph=g2)
        t0 = time.time()
        text_lengths = get_text_lengths(L)
        K, V = encode_text(hp, L, g1, sess)
        Y, lengths, alignments = synth_codedtext2mel_fast(hp, K, V, text_lengths, g1)
        print('Text2Mel generating... %.2f seconds' % (time.time() - t0))
        t0 = time.time()
        Z = synth_mel2mag(hp, Y, g2, sess)
        print('Mel2Mag generating... %.2f seconds' % (time.time() - t0))
        if np.isnan(Z).any():
            Z = np.nan_to_num(Z)
        outdir = os.path.join(topoutdir or hp.sampledir, 't2m%s_ssrn%s' % (t2m_epoch, ssrn_epoch))
        os.makedirs(outdir, exist_ok=True)
        print("File |  CDP | Ain")
        for i in range(len(Z)):
            trimmed_alignment = alignments[i, :text_lengths[i], :lengths[i]]
            np.save(os.path.join(outdir, bases[i] + '_alignment.npy'), trimmed_alignment)
            if trimmed_alignment.size:
                APin, _APout = getAP(trimmed_alignment)
                print("%s | %.2f | %.2f" % (bases[i], getCDP(trimmed_alignment), APin))
        print("Generating wav files, will save to following dir: %s" % (outdir))
        for i, mag in enumerate(Z):
            mag = mag[:lengths[i] * hp.r, :]                       # trim to generated length
            outfile = os.path.join(outdir, bases[i] + '.wav')
            if vocode:
                vocoder.synth_wave(hp, mag, outfile)
            else:
                np.save(outfile.replace('.wav', '.npy'), mag)
    return outdir, bases, lengths


def main_work():
    import os
    import re
    from argparse import ArgumentParser
    from .configuration import load_config
    a = ArgumentParser()
    a.add_argument('-c', dest='config', required=True, type=str)
    a.add_argument('-speaker', default='', type=str)
    a.add_argument('-N', dest='num_sentences', default=0, type=int)
    a.add_argument('-ncores', type=int, default=1, help='accepted for compatibility: Griffin-Lim runs on the GPU')
    a.add_argument('-odir', type=str, default='', help='Alternative place to put output samples')
    a.add_argument('-t2m_epoch', default=-1, type=int, help='Default: use latest (-1)')
    a.add_argument('-ssrn_epoch', default=-1, type=int, help='Default: use latest (-1)')
    a.add_argument('-max_N', default=-1, type=int, help='Default: use max_N from config')
    a.add_argument('-max_T', default=-1, type=int, help='Default: use max_T from config')
    a.add_argument('-tr', default='', type=str, help='Default:use test_transcript from config')
    opts = a.parse_args()
    hp = load_config(opts.config)
    if opts.max_N != -1:
        hp.max_N = opts.max_N
    if opts.max_T != -1:
        hp.max_T = opts.max_T
    if opts.tr != '':
        hp.test_transcript = opts.tr
    print("max_N=" + str(hp.max_N))
    print("max_T=" + str(hp.max_T))
    print("test_transcript=" + str(hp.test_transcript))
    outdir = opts.odir
    if outdir:
        outdir = os.path.join(outdir, re.sub(r'\.[^\.]+\Z', '', os.path.split(opts.config)[1]))
    synthesize(hp, speaker_id=opts.speaker, num_sentences=opts.num_sentences, ncores=opts.ncores, topoutdir=outdir,
               t2m_epoch=opts.t2m_epoch, ssrn_epoch=opts.ssrn_epoch)


if __name__ == "__main__":
    main_work()
