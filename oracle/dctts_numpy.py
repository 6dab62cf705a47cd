"""numpy fp64 restatement of the dc_tts hot path -- the arbiter oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
reference cannot be executed here (TensorFlow 1.12 / Python 2.7), and it holds
no golden vectors for this path; every function cites the reference lines it
restates and the TF-1.12 op semantics it assumes.

All tensors are channels-last `[B, time, C]` (TF NWC), parameters are a dict
`name -> ndarray` keyed by the reference's checkpoint variable names
(`train.py:194`):  `<scope>/conv1d/{kernel[k,Cin,Cout],bias}`,
`<scope>/normalize/{beta,gamma}`, `<scope>/{H1,H2}/{beta,gamma}`,
`<scope>/conv2d_transpose/{kernel[1,3,Cout,Cin],bias}`,
`<scope>/lookup_table`.

Loops/einsum only; no torch.  Inputs are promoted to float64.
"""
import numpy as np

LN_EPS = 1e-12          # tf.contrib.layers.layer_norm -> batch_normalization variance_epsilon
MASK_VALUE = float(-2 ** 32 + 1)   # networks.py:312


# --------------------------------------------------------------------------- modules.py

def embed(P, inputs, scope, zero_pad=True):
    """modules.py:15-44.  Row 0 of the table is replaced by zeros (:38-40)."""
    table = np.asarray(P[scope + "/lookup_table"], np.float64)
    if zero_pad:
        table = np.concatenate([np.zeros((1, table.shape[1])), table[1:]], 0)
    return table[np.asarray(inputs)]


def normalize(P, x, scope, normtype="layer"):
    """modules.py:47-75.  'layer' = tf.contrib.layers.layer_norm(begin_norm_axis=-1):
    biased variance over the last axis, eps 1e-12, gamma/beta of shape [C]."""
    assert normtype in (None, "layer")
    if normtype is None:
        return x
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    xh = (x - mu) / np.sqrt(var + LN_EPS)
    return xh * np.asarray(P[scope + "/gamma"], np.float64) + np.asarray(P[scope + "/beta"], np.float64)


def _dilated_conv(x, kernel, bias, rate, padding):
    """tf.layers.conv1d at modules.py:134-136 / :189-193.
    kernel [k, Cin, Cout]; 'same' pads (k-1)*rate total with left = total//2;
    'causal' left-pads (k-1)*rate explicitly and runs VALID (modules.py:123-127)."""
    k = kernel.shape[0]
    B, L, _ = x.shape
    total = (k - 1) * rate
    padding = padding.lower()
    if padding == "causal":
        left = total
    elif padding == "same":
        left = total // 2
    else:
        raise ValueError(padding)
    xp = np.zeros((B, L + total, x.shape[2]))
    xp[:, left:left + L] = x
    out = np.zeros((B, L, kernel.shape[2]))
    for j in range(k):
        out += np.einsum("blc,cd->bld", xp[:, j * rate:j * rate + L], kernel[j])
    return out + bias


def conv1d(P, x, scope, size=1, rate=1, padding="SAME", activation_fn=None, normtype="layer"):
    """modules.py:91-146 with training=False (dropout is identity)."""
    kernel = np.asarray(P[scope + "/conv1d/kernel"], np.float64)
    assert kernel.shape[0] == size
    bias = np.asarray(P[scope + "/conv1d/bias"], np.float64)
    t = _dilated_conv(x, kernel, bias, rate, padding)
    t = normalize(P, t, scope + "/normalize", normtype)
    if activation_fn == "relu":
        t = np.maximum(t, 0.0)
    elif activation_fn is not None:
        raise ValueError(activation_fn)
    return t


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def hc(P, x, scope, size=1, rate=1, padding="SAME", normtype="layer"):
    """modules.py:148-207 highway conv: conv to 2C, split, LN(H1), LN(H2) separately,
    sigmoid gate, H1*H2 + (1-H1)*inputs (un-padded input)."""
    kernel = np.asarray(P[scope + "/conv1d/kernel"], np.float64)
    assert kernel.shape[0] == size
    bias = np.asarray(P[scope + "/conv1d/bias"], np.float64)
    t = _dilated_conv(x, kernel, bias, rate, padding)
    C = t.shape[-1] // 2
    H1, H2 = t[..., :C], t[..., C:]
    H1 = normalize(P, H1, scope + "/H1", normtype)
    H2 = normalize(P, H2, scope + "/H2", normtype)
    H1 = sigmoid(H1)
    return H1 * H2 + (1.0 - H1) * x


def conv1d_transpose(P, x, scope, normtype="layer"):
    """modules.py:209-258.  conv2d_transpose(kernel (1,3), strides (1,2), 'same') with kernel
    [1,3,Cout,Cin] is the gradient of a stride-2 SAME conv (pad_left 0, pad_right 1), hence
    out[2i] = W[0].x[i] + W[2].x[i-1],  out[2i+1] = W[1].x[i].
    LN is applied with the function's default normtype='layer' -- the caller
    (networks.py:483-486) never passes hp.norm."""
    W = np.asarray(P[scope + "/conv2d_transpose/kernel"], np.float64)[0]   # [3, Cout, Cin]
    bias = np.asarray(P[scope + "/conv2d_transpose/bias"], np.float64)
    B, L, _ = x.shape
    out = np.zeros((B, 2 * L, W.shape[1]))
    xm1 = np.zeros_like(x)
    xm1[:, 1:] = x[:, :-1]
    out[:, 0::2] = np.einsum("blc,dc->bld", x, W[0]) + np.einsum("blc,dc->bld", xm1, W[2])
    out[:, 1::2] = np.einsum("blc,dc->bld", x, W[1])
    out = out + bias
    return normalize(P, out, scope + "/normalize", normtype)


# --------------------------------------------------------------------------- networks.py

def TextEnc(hp, P, L, prefix="Text2Mel/TextEnc"):
    """networks.py:121-212 (default LJ path: no speaker codes)."""
    i = 1
    t = embed(P, L, "%s/embed_%d" % (prefix, i)); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), activation_fn="relu", normtype=hp.norm); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), normtype=hp.norm); i += 1
    for _ in range(2):
        for j in range(4):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3 ** j, normtype=hp.norm); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=1, normtype=hp.norm); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=1, rate=1, normtype=hp.norm); i += 1
    d = t.shape[-1] // 2
    return t[..., :d], t[..., d:]


def AudioEnc(hp, P, S, prefix="Text2Mel/AudioEnc"):
    """networks.py:214-284: every layer CAUSAL."""
    i = 1
    t = conv1d(P, S, "%s/C_%d" % (prefix, i), padding="CAUSAL", activation_fn="relu", normtype=hp.norm); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), padding="CAUSAL", activation_fn="relu", normtype=hp.norm); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), padding="CAUSAL", normtype=hp.norm); i += 1
    for _ in range(2):
        for j in range(4):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3 ** j, padding="CAUSAL", normtype=hp.norm); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3, padding="CAUSAL", normtype=hp.norm); i += 1
    return t


def Attention(hp, Q, K, V, monotonic_attention=False, prev_max_attentions=None, text_lengths=None):
    """networks.py:286-325.  Returns (R', alignments [B,N,T], max_attentions [B,T])."""
    d = float(hp.d)
    A = np.einsum("btd,bnd->btn", Q, K) * (1.0 / np.sqrt(d))
    if monotonic_attention:
        B, T, N = A.shape
        n = np.arange(N)[None, :]
        if not getattr(hp, "turn_off_monotonic_for_synthesis", False):
            assert N == hp.max_N and T == hp.max_T          # sequence_mask(.., hp.max_N), tile(.., hp.max_T)
            prev = np.asarray(prev_max_attentions).reshape(B, 1)
            key_masks = n < prev                                               # :304
            reverse = (n < (hp.max_N - hp.attention_win_size - prev))[:, ::-1]  # :305
            masks = key_masks | reverse
        else:
            tl = np.asarray(text_lengths).reshape(B, 1)
            masks = (n < (hp.max_N - tl))[:, ::-1]
        masks = np.broadcast_to(masks[:, None, :], A.shape)                    # same mask on every t (:311)
        A = np.where(masks, MASK_VALUE, A)
    A = A - A.max(-1, keepdims=True)
    A = np.exp(A)
    A = A / A.sum(-1, keepdims=True)
    max_attentions = A.argmax(-1)
    R = np.einsum("btn,bnd->btd", A, V)
    if getattr(hp, "concatenate_query", True):
        R = np.concatenate([R, Q], -1)
    alignments = A.transpose(0, 2, 1)
    return R, alignments, max_attentions


def AudioDec(hp, P, R, prefix="Text2Mel/AudioDec"):
    """networks.py:360-435.  Returns (logits, Y)."""
    i = 1
    t = conv1d(P, R, "%s/C_%d" % (prefix, i), padding="CAUSAL", normtype=hp.norm); i += 1
    for j in range(4):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3 ** j, padding="CAUSAL", normtype=hp.norm); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=1, padding="CAUSAL", normtype=hp.norm); i += 1
    for _ in range(3):
        t = conv1d(P, t, "%s/C_%d" % (prefix, i), padding="CAUSAL", activation_fn="relu", normtype=hp.norm); i += 1
    logits = conv1d(P, t, "%s/C_%d" % (prefix, i), padding="CAUSAL", normtype=hp.norm); i += 1
    Y = sigmoid(logits) if getattr(hp, "squash_output_t2m", True) else logits
    return logits, Y


def SSRN(hp, P, Y, prefix="SSRN"):
    """networks.py:437-537.  Returns (logits, Z)."""
    i = 1
    t = conv1d(P, Y, "%s/C_%d" % (prefix, i), normtype=hp.norm); i += 1
    for j in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3 ** j, normtype=hp.norm); i += 1
    n_transposes = {4: 2, 8: 3}[hp.r]
    for _ in range(n_transposes):
        t = conv1d_transpose(P, t, "%s/D_%d" % (prefix, i)); i += 1
        for j in range(2):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=3 ** j, normtype=hp.norm); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), normtype=hp.norm); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), size=3, rate=1, normtype=hp.norm); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), normtype=hp.norm); i += 1
    for _ in range(2):
        t = conv1d(P, t, "%s/C_%d" % (prefix, i), activation_fn="relu", normtype=hp.norm); i += 1
    logits = conv1d(P, t, "%s/C_%d" % (prefix, i), normtype=hp.norm)
    Z = sigmoid(logits) if getattr(hp, "squash_output_ssrn", True) else logits
    return logits, Z


# --------------------------------------------------------------------------- utils.py / architectures.py

def get_attention_guide(xdim, ydim, g=0.2):
    """utils.py:155-161 (float32 table of 1-exp(-(t/T-n/N)^2/2g^2))."""
    n = np.arange(xdim, dtype=np.float64)[:, None] / float(xdim)
    t = np.arange(ydim, dtype=np.float64)[None, :] / float(ydim)
    return (1.0 - np.exp(-(t - n) ** 2 / (2 * g * g))).astype(np.float32)


def learning_rate_decay(init_lr, global_step, warmup_steps=4000.0):
    """utils.py:167-170 Noam schedule on step+1."""
    step = float(global_step + 1)
    return init_lr * warmup_steps ** 0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)


def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    return np.maximum(logits, 0) - logits * labels + np.log1p(np.exp(-np.abs(logits)))


def text2mel_forward(hp, P, L, mels, mode="train", prev_max_attentions=None, K=None, V=None):
    """architectures.py:188-239 (DCTTS_standard encoder/history).  K,V can be fed (synthesize.py:172)."""
    mels = np.asarray(mels, np.float64)
    S = np.concatenate([np.zeros_like(mels[:, :1]), mels[:, :-1]], 1)        # :191
    if K is None:
        K, V = TextEnc(hp, P, L)
    Q = AudioEnc(hp, P, S)
    mono = (mode == "synthesize")
    R, alignments, max_att = Attention(hp, Q, np.asarray(K, np.float64), np.asarray(V, np.float64),
                                       monotonic_attention=mono, prev_max_attentions=prev_max_attentions,
                                       text_lengths=getattr(hp, "text_lengths", None))
    logits, Y = AudioDec(hp, P, R)
    return dict(K=K, V=V, Q=Q, R=R, alignments=alignments, max_attentions=max_att, Y_logits=logits, Y=Y)


def _loss_weights(hp, which):
    lw = getattr(hp, "loss_weights", None)
    if which == "t2m":
        if lw and "t2m" in lw:
            w = lw["t2m"]
            return w["L1"], w["binary_divergence"], w["attention"], w["L2"]
        return hp.lw_mel, hp.lw_bd1, hp.lw_att, getattr(hp, "lw_t2m_l2", 0.0)
    if lw and "ssrn" in lw:
        w = lw["ssrn"]
        return w["L1"], w["binary_divergence"], w["L2"]
    return hp.lw_mag, hp.lw_bd2, getattr(hp, "lw_ssrn_l2", 0.0)


def _pad_crop(x, value, max_N, max_T):
    """`tf.pad(x, [(0,0),(0,max_N),(0,max_T)], constant_values=value)[:, :max_N, :max_T]` (architectures.py:261-263)."""
    B, n, t = x.shape
    out = np.full((B, n + max_N, t + max_T), float(value))
    out[:, :n, :t] = x
    return out[:, :max_N, :max_T]


def text2mel_loss(hp, out, mels, guide=None, gts=None):
    """architectures.py:241-355 (lw_cdp=ain=aout=0).  Returns loss_components [loss, L1, BD, att, L2].
    gts: the batch's per-utterance guides / forced-alignment targets [B, Ng, Tg] (hp.attention_guide_dir, :57-58), zero
    padded within the batch like the reference's dynamic_pad queue; None = global guide (:60).  hp.attention_guide_fa
    selects the MSE branch (:271-280)."""
    mels = np.asarray(mels, np.float64)
    Y, logits, A = out["Y"], out["Y_logits"], out["alignments"]
    loss_l2 = ((Y - mels) ** 2).mean()
    loss_mels = np.abs(Y - mels).mean()
    loss_bd1 = sigmoid_ce(logits, mels).mean() if getattr(hp, "squash_output_t2m", True) else 0.0
    if guide is None:
        guide = get_attention_guide(hp.max_N, hp.max_T, hp.g)
    # pad with -1 to [max_N,max_T] then crop (:262); mask = (A != -1)
    Bn, Nb, Tb = A.shape
    Ap = -np.ones((Bn, Nb + hp.max_N, Tb + hp.max_T))
    Ap[:, :Nb, :Tb] = A
    Ap = Ap[:, :hp.max_N, :hp.max_T]
    mask = (Ap != -1).astype(np.float64)
    if gts is not None and getattr(hp, "attention_guide_fa", False):            # :271-280
        A0 = _pad_crop(A, 0.0, hp.max_N, hp.max_T)
        G0 = _pad_crop(np.asarray(gts, np.float64), 0.0, hp.max_N, hp.max_T)
        loss_att = ((A0 - G0) ** 2).sum() / mask.sum()
    elif gts is not None:                                                        # :262-268 with :263's padding by 1.0
        G1 = _pad_crop(np.asarray(gts, np.float64), 1.0, hp.max_N, hp.max_T)
        loss_att = (np.abs(Ap * G1) * mask).sum() / mask.sum()
    else:
        loss_att = (np.abs(Ap * np.asarray(guide, np.float64)[None]) * mask).sum() / mask.sum()
    w1, wbd, watt, w2 = _loss_weights(hp, "t2m")
    loss = w1 * loss_mels + wbd * loss_bd1 + watt * loss_att + w2 * loss_l2
    lw_cdp, lw_ain, lw_aout = (getattr(hp, n, 0.0) for n in ("lw_cdp", "lw_ain", "lw_aout"))
    if lw_cdp != 0.0 or lw_ain != 0.0 or lw_aout != 0.0:                        # :283-321, :333-355
        cdp, ain, aout = attention_confidence_terms(A)
        lw = getattr(hp, "loss_weights", None)
        if not (lw and "t2m" in lw):        # the loss_weights dict branch (:325-331) does not add these terms
            loss = loss + (lw_cdp * cdp if lw_cdp != 0.0 else 0.0) + (lw_ain * ain if lw_ain != 0.0 else 0.0) \
                + (lw_aout * aout if lw_aout != 0.0 else 0.0)
        return [loss, loss_mels, loss_bd1, loss_att, loss_l2, cdp, ain, aout]
    return [loss, loss_mels, loss_bd1, loss_att, loss_l2]


def _plogp(P):
    return np.where(P != 0, P * np.log(np.where(P != 0, P, 1.0)), 0.0)


def attention_confidence_terms(A):
    """architectures.py:283-321 on alignments [B, N_b, T_b]: coverage deviation penalty and the entropies of the attention
    per input symbol (Ain, normalised over frames) and per output frame (Aout, normalised over symbols)."""
    B, N, T = A.shape
    att_per_input = A.sum(2)
    cdp = np.log(1.0 + (1.0 - att_per_input) ** 2).sum() / (B * N)
    with np.errstate(invalid="ignore", divide="ignore"):
        Pin = A / A.sum(2, keepdims=True)
    Pin = np.where(np.isnan(Pin), 0.0, Pin)                                     # :301
    ain = -_plogp(Pin).sum() / (B * N) / np.log(T)
    Pout = A / A.sum(1, keepdims=True)
    aout = -_plogp(Pout).sum() / (B * T) / np.log(N)
    return cdp, ain, aout


def ssrn_loss(hp, logits, Z, mags):
    """architectures.py:144-173.  loss_components [loss, L1, BD, L2]."""
    mags = np.asarray(mags, np.float64)
    loss_l2 = ((Z - mags) ** 2).mean()
    loss_mags = np.abs(Z - mags).mean()
    loss_bd2 = sigmoid_ce(logits, mags).mean() if getattr(hp, "squash_output_ssrn", True) else 0.0
    w1, wbd, w2 = _loss_weights(hp, "ssrn")
    return [w1 * loss_mags + wbd * loss_bd2 + w2 * loss_l2, loss_mags, loss_bd2, loss_l2]


def adam_step(p, m, v, g, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, clip=1.0):
    """architectures.py:110-128: clip_by_value(grad,-1,1) then tf.train.AdamOptimizer:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).  t is the 1-based step."""
    g = np.clip(g, -clip, clip)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


# --------------------------------------------------------------------------- synthesize.py

def synth_codedtext2mel(hp, P, K, V, ends):
    """synthesize.py:150-230: autoregressive loop re-running the whole graph per frame.
    Returns (Y [B,max_T,n_mels], t_ends list, alignments [B,max_N,max_T])."""
    B = len(K)
    Y = np.zeros((B, hp.max_T, hp.n_mels))
    alignments = np.zeros((B, hp.max_N, hp.max_T))
    prev = np.zeros((B,), np.int64)
    ends = np.asarray(ends)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * hp.max_T
    for j in range(hp.max_T):
        out = text2mel_forward(hp, P, None, Y, mode="synthesize", prev_max_attentions=prev, K=K, V=V)
        Y[:, j] = out["Y"][:, j]
        alignments[:, :, j] = out["alignments"][:, :, j]
        prev = out["max_attentions"][:, j]
        endcounts += (prev >= ends)
        for i in range(B):
            if t_ends[i] == hp.max_T and endcounts[i] >= 1:
                t_ends[i] = j
        if (t_ends < hp.max_T).all():
            break
    return Y, t_ends.tolist(), alignments
