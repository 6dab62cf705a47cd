"""torch-CPU restatement of the research variants that hang off the same operator surface (TEST INFRASTRUCTURE ONLY;
PARITY UNPINNED like the rest of oracle/): speaker embeddings at the five positions of hp.multispeaker, MerlinTextEnc /
LinearTransformLabels / the label-only text encoders, FixedAttention with external durations, BabblerGraph, and the
per-speaker channel gates of modules.learn_channel_contributions (modules.py:78-88, 141-144, 200-201).
Cites networks.py:15-119 (MerlinTextEnc), :121-212 (TextEnc), :214-284 (AudioEnc), :327-358 (FixedAttention),
:360-435 (AudioDec), :437-537 (SSRN), :540-560 (LinearTransformLabels), architectures.py:183-239, 380-432.
The layer counter `i` of every network is advanced exactly like the reference's (speaker embeddings consume an index)."""
import torch
import torch.nn.functional as F

from . import dctts_torch as ot


def _ms(hp):
    return getattr(hp, "multispeaker", [])


def speaker_cat(hp, P, t, scope, speaker_codes):
    """tf.tile(speaker_codes, [1, L]) -> embed (zero-padded table, modules.py:38-40) -> tf.concat((tensor, reps), -1)."""
    B, L, _ = t.shape
    ids = torch.as_tensor(speaker_codes, dtype=torch.long).reshape(B, 1).expand(B, L)
    return torch.cat([t, ot.embed(P, ids, scope)], -1)


def _lcc(hp):
    return 'learn_channel_contributions' in _ms(hp)


def lcc_gate(P, scope, codes):
    """modules.learn_channel_contributions (modules.py:78-88): sigmoid(embed(codes [B, 1], scope="lcc_embed")) -> [B, 1, C]."""
    ids = torch.as_tensor(codes, dtype=torch.long).reshape(-1, 1)
    return torch.sigmoid(ot.embed(P, ids, scope + "/lcc_embed"))


def conv1d(hp, P, x, scope, codes=None, gated=False, act=None, **kw):
    """modules.conv1d with the optional channel gates after dropout (modules.py:141-144)."""
    t = ot.conv1d(P, x, scope, act=act, **kw)
    if gated and _lcc(hp):
        t = lcc_gate(P, scope, codes) * t
    return t


def hc(hp, P, x, scope, size, rate, codes=None, gated=True, padding="SAME", normtype="layer", dropout_rate=0.0, training=False,
       gen=None):
    """modules.hc with the channel gates on the transformation connection H2 only (modules.py:194-205)."""
    if not (gated and _lcc(hp)):
        return ot.hc(P, x, scope, size, rate, padding=padding, normtype=normtype, dropout_rate=dropout_rate, training=training,
                     gen=gen)
    t = ot._conv(x, P[scope + "/conv1d/kernel"], P[scope + "/conv1d/bias"], rate, padding)
    H1, H2 = t.chunk(2, -1)
    H1 = torch.sigmoid(ot.layer_norm(P, H1, scope + "/H1", normtype))
    H2 = ot.layer_norm(P, H2, scope + "/H2", normtype)
    H2 = lcc_gate(P, scope, codes) * H2
    return ot._drop(H1 * H2 + (1.0 - H1) * x, dropout_rate, training, gen)


def _text_body(hp, P, t, i, prefix, speaker_codes, kw):
    c = speaker_codes
    t = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, act="relu", **kw); i += 1
    t = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, **kw); i += 1
    for _ in range(2):
        for j in range(4):
            t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, c, **kw); i += 1
    for _ in range(2):
        t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 1, c, **kw); i += 1
    if 'text_encoder_towards_end' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
        t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), act="relu", **kw); i += 1
    for _ in range(2):
        t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 1, 1, c, **kw); i += 1
    return t.chunk(2, -1)


def TextEnc(hp, P, L, speaker_codes=None, training=False, gen=None, prefix="Text2Mel/TextEnc"):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i = 1
    t = ot.embed(P, L, "%s/embed_%d" % (prefix, i)); i += 1
    if 'text_encoder_input' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
    return _text_body(hp, P, t, i, prefix, speaker_codes, kw)


def LinearTransformLabels(hp, P, labels, prefix, i=1, training=False, gen=None):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    return ot.conv1d(P, labels, "%s/C_%d" % (prefix, i), **kw)


def MerlinTextEnc(hp, P, L, labels, speaker_codes=None, training=False, gen=None, prefix="Text2Mel/MerlinTextEnc"):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i = 1
    t = LinearTransformLabels(hp, P, labels, prefix, 1, training, gen); i += 1
    if hp.MerlinTextEncWithPhoneEmbedding:
        t = torch.cat([t, ot.embed(P, L, "%s/embed_%d" % (prefix, i))], -1); i += 1
    if 'text_encoder_input' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
    return _text_body(hp, P, t, i, prefix, speaker_codes, kw)


def AudioEnc(hp, P, S, speaker_codes=None, training=False, gen=None, prefix="Text2Mel/AudioEnc"):
    kw = dict(padding="CAUSAL", normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i, c = 1, speaker_codes
    t = conv1d(hp, P, S, "%s/C_%d" % (prefix, i), c, True, act="relu", **kw); i += 1
    if 'audio_encoder_input' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
        t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **dict(kw, padding="SAME")); i += 1
    t = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, act="relu", **kw); i += 1
    t = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, **kw); i += 1
    for _ in range(2):
        for j in range(4):
            t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, c, **kw); i += 1
    for _ in range(2):
        t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 3, c, **kw); i += 1
    return t


def FixedAttention(hp, duration_matrix, Q, V):
    mx = duration_matrix.argmax(-1)
    R = torch.matmul(duration_matrix, V)
    if getattr(hp, "concatenate_query", True):
        R = torch.cat([R, Q], -1)
    return R, duration_matrix.transpose(1, 2), mx


def AudioDec(hp, P, R, speaker_codes=None, training=False, gen=None, prefix="Text2Mel/AudioDec"):
    kw = dict(padding="CAUSAL", normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i = 1
    t = ot.conv1d(P, R, "%s/C_%d" % (prefix, i), **kw); i += 1
    if 'audio_decoder_input' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
        t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **dict(kw, padding="SAME")); i += 1
    c = speaker_codes
    for j in range(4):
        t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, c, **kw); i += 1
    for _ in range(2):
        t = hc(hp, P, t, "%s/HC_%d" % (prefix, i), 3, 1, c, **kw); i += 1
    for _ in range(3):
        t = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, act="relu", **kw); i += 1
    logits = conv1d(hp, P, t, "%s/C_%d" % (prefix, i), c, True, **kw)
    Y = torch.sigmoid(logits) if getattr(hp, "squash_output_t2m", True) else logits
    return logits, Y


def SSRN(hp, P, Y, speaker_codes=None, training=False, gen=None, prefix="SSRN"):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i = 1
    t = ot.conv1d(P, Y, "%s/C_%d" % (prefix, i), **kw); i += 1
    if 'ssrn_input' in _ms(hp):
        t = speaker_cat(hp, P, t, "%s/embed_%d" % (prefix, i), speaker_codes); i += 1
        t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **kw); i += 1
    for j in range(2):
        t = ot.hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    for _ in range({4: 2, 8: 3}[hp.r]):
        t = ot.conv1d_transpose(P, t, "%s/D_%d" % (prefix, i), hp.dropout_rate, training, gen); i += 1
        for j in range(2):
            t = ot.hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **kw); i += 1
    for _ in range(2):
        t = ot.hc(P, t, "%s/HC_%d" % (prefix, i), 3, 1, **kw); i += 1
    t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **kw); i += 1
    for _ in range(2):
        t = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), act="relu", **kw); i += 1
    logits = ot.conv1d(P, t, "%s/C_%d" % (prefix, i), **kw)
    Z = torch.sigmoid(logits) if getattr(hp, "squash_output_ssrn", True) else logits
    return logits, Z


def text2mel_forward(hp, P, L, mels, mode="train", prev_max_attentions=None, speakers=None, durations=None, labels=None,
                     gen=None):
    """architectures.py:188-239 with every branch of text_encoder_type / use_external_durations / multispeaker."""
    training = mode == "train"
    S = torch.cat([torch.zeros_like(mels[:, :1]), mels[:, :-1]], 1)
    enc = hp.text_encoder_type
    if enc == 'none':
        K = V = labels
    elif enc == 'minimal_feedforward':
        K = V = LinearTransformLabels(hp, P, labels, "Text2Mel", 1, training, gen)
    elif enc == 'MerlinTextEnc':
        K, V = MerlinTextEnc(hp, P, L, labels, speakers, training, gen)
    else:
        K, V = TextEnc(hp, P, L, speakers, training, gen)
    Q = AudioEnc(hp, P, S, speakers, training, gen)
    if hp.use_external_durations:
        R, ali, mx = FixedAttention(hp, durations, Q, V)
    else:
        R, ali, mx = ot.Attention(hp, Q, K, V, mode == "synthesize", prev_max_attentions)
    logits, Y = AudioDec(hp, P, R, speakers, training, gen)
    return dict(K=K, V=V, Q=Q, R=R, alignments=ali, max_attentions=mx, Y_logits=logits, Y=Y)


def text2mel_train_step(hp, P, opt, L, mels, gen=None, **inputs):
    for p in P.values():
        p.grad = None
    out = text2mel_forward(hp, P, L, mels, "train", gen=gen, **inputs)
    comps = ot.text2mel_loss(hp, out, mels)
    comps[0].backward()
    grads = {k: p.grad for k, p in P.items() if p.grad is not None}
    opt.step(grads)
    return [float(c.detach()) for c in comps], grads


def ssrn_train_step(hp, P, opt, mels, mags, speakers=None, gen=None):
    for p in P.values():
        p.grad = None
    logits, Z = SSRN(hp, P, mels, speakers, True, gen)
    comps = ot.ssrn_loss(hp, logits, Z, mags)
    comps[0].backward()
    grads = {k: p.grad for k, p in P.items() if p.grad is not None}
    opt.step(grads)
    return [float(c.detach()) for c in comps], grads


def babbler_forward(hp, P, mels, training=False, gen=None):
    """architectures.py:394-410: R = concat(zeros_like(Q), Q)."""
    S = torch.cat([torch.zeros_like(mels[:, :1]), mels[:, :-1]], 1)
    Q = ot.AudioEnc(hp, P, S, training, gen)
    R = torch.cat([torch.zeros_like(Q), Q], -1)
    logits, Y = ot.AudioDec(hp, P, R, training, gen)
    return dict(Q=Q, R=R, Y_logits=logits, Y=Y)


def babbler_train_step(hp, P, opt, mels, gen=None):
    """architectures.py:412-424: loss = w_L1 * mean|Y - mels| + w_bd * mean sigmoid-CE(logits, mels)."""
    for p in P.values():
        p.grad = None
    out = babbler_forward(hp, P, mels, True, gen)
    l1 = (out["Y"] - mels).abs().mean()
    bd = F.binary_cross_entropy_with_logits(out["Y_logits"], mels)
    w = hp.loss_weights['babbler']
    loss = w['L1'] * l1 + w['binary_divergence'] * bd
    loss.backward()
    grads = {k: p.grad for k, p in P.items() if p.grad is not None}
    opt.step(grads)
    return [float(loss.detach()), float(l1.detach()), float(bd.detach())], grads


def synth_babble(hp, P, nsamples):
    """synthesize.py:134-148."""
    Y = torch.zeros(nsamples, hp.max_T, hp.n_mels, dtype=next(iter(P.values())).dtype)
    with torch.no_grad():
        for j in range(hp.max_T):
            Y[:, j] = babbler_forward(hp, P, Y)["Y"][:, j]
    return Y
