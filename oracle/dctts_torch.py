"""torch-CPU restatement of the dc_tts hot path with autograd (TEST INFRASTRUCTURE ONLY).

Second, independent oracle (uses torch's conv kernels instead of einsum): fp32 = stand-in for the
reference-precision CPU path and the timed CPU baseline; fp64 autograd = gradient goldens.
PARITY UNPINNED (see oracle/__init__.py).  Cites: modules.py, networks.py, architectures.py,
utils.py:155-170 of the reference.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-12
MASK_VALUE = float(-2 ** 32 + 1)


def to_torch(P, dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in P.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def _drop(x, rate, training, gen):
    if not training or rate <= 0:
        return x
    keep = 1.0 - rate
    mask = (torch.rand(x.shape, generator=gen, dtype=x.dtype) < keep).to(x.dtype)
    return x * mask / keep


def embed(P, ids, scope):
    table = P[scope + "/lookup_table"]
    table = torch.cat([torch.zeros_like(table[:1]), table[1:]], 0)           # modules.py:38-40
    return table[torch.as_tensor(ids, dtype=torch.long)]


def layer_norm(P, x, scope, normtype="layer"):
    if normtype is None:
        return x
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * P[scope + "/gamma"] + P[scope + "/beta"]


def _conv(x, kernel, bias, rate, padding):
    k = kernel.shape[0]
    total = (k - 1) * rate
    left = total if padding.lower() == "causal" else total // 2
    xt = F.pad(x.transpose(1, 2), (left, total - left))                       # [B,C,L+total]
    w = kernel.permute(2, 1, 0)                                               # [Cout,Cin,k]
    return F.conv1d(xt, w, bias, dilation=rate).transpose(1, 2)


def conv1d(P, x, scope, size=1, rate=1, padding="SAME", act=None, normtype="layer",
           dropout_rate=0.0, training=False, gen=None):
    t = _conv(x, P[scope + "/conv1d/kernel"], P[scope + "/conv1d/bias"], rate, padding)
    t = layer_norm(P, t, scope + "/normalize", normtype)
    if act == "relu":
        t = torch.relu(t)
    return _drop(t, dropout_rate, training, gen)


def hc(P, x, scope, size=1, rate=1, padding="SAME", normtype="layer", dropout_rate=0.0, training=False, gen=None):
    t = _conv(x, P[scope + "/conv1d/kernel"], P[scope + "/conv1d/bias"], rate, padding)
    H1, H2 = t.chunk(2, -1)
    H1 = torch.sigmoid(layer_norm(P, H1, scope + "/H1", normtype))
    H2 = layer_norm(P, H2, scope + "/H2", normtype)
    return _drop(H1 * H2 + (1.0 - H1) * x, dropout_rate, training, gen)


def conv1d_transpose(P, x, scope, dropout_rate=0.0, training=False, gen=None):
    W = P[scope + "/conv2d_transpose/kernel"][0]                              # [3,Cout,Cin]
    w = W.permute(2, 1, 0)                                                    # conv_transpose1d weight [Cin,Cout,k]
    L = x.shape[1]
    t = F.conv_transpose1d(x.transpose(1, 2), w, P[scope + "/conv2d_transpose/bias"], stride=2)[..., :2 * L]
    t = layer_norm(P, t.transpose(1, 2), scope + "/normalize", "layer")
    return _drop(t, dropout_rate, training, gen)


def TextEnc(hp, P, L, training=False, gen=None, prefix="Text2Mel/TextEnc"):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    t = embed(P, L, prefix + "/embed_1")
    t = conv1d(P, t, prefix + "/C_2", act="relu", **kw)
    t = conv1d(P, t, prefix + "/C_3", **kw)
    i = 4
    for _ in range(2):
        for j in range(4):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 1, **kw); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 1, 1, **kw); i += 1
    return t.chunk(2, -1)


def AudioEnc(hp, P, S, training=False, gen=None, prefix="Text2Mel/AudioEnc"):
    kw = dict(padding="CAUSAL", normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    t = conv1d(P, S, prefix + "/C_1", act="relu", **kw)
    t = conv1d(P, t, prefix + "/C_2", act="relu", **kw)
    t = conv1d(P, t, prefix + "/C_3", **kw)
    i = 4
    for _ in range(2):
        for j in range(4):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3, **kw); i += 1
    return t


def Attention(hp, Q, K, V, monotonic_attention=False, prev_max_attentions=None):
    A = torch.matmul(Q, K.transpose(1, 2)) * (1.0 / math.sqrt(float(hp.d)))
    if monotonic_attention:
        B, T, N = A.shape
        assert N == hp.max_N and T == hp.max_T
        n = torch.arange(N)[None, :]
        prev = torch.as_tensor(prev_max_attentions).reshape(B, 1)
        key_masks = n < prev
        reverse = torch.flip(n < (hp.max_N - hp.attention_win_size - prev), dims=[1])
        masks = (key_masks | reverse)[:, None, :].expand_as(A)
        A = torch.where(masks, torch.full_like(A, MASK_VALUE), A)
    A = torch.softmax(A, -1)
    max_att = A.argmax(-1)
    R = torch.matmul(A, V)
    if getattr(hp, "concatenate_query", True):
        R = torch.cat([R, Q], -1)
    return R, A.transpose(1, 2), max_att


def AudioDec(hp, P, R, training=False, gen=None, prefix="Text2Mel/AudioDec"):
    kw = dict(padding="CAUSAL", normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    t = conv1d(P, R, prefix + "/C_1", **kw)
    i = 2
    for j in range(4):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 1, **kw); i += 1
    for _ in range(3):
        t = conv1d(P, t, "%s/C_%d" % (prefix, i), act="relu", **kw); i += 1
    logits = conv1d(P, t, "%s/C_%d" % (prefix, i), **kw)
    Y = torch.sigmoid(logits) if getattr(hp, "squash_output_t2m", True) else logits
    return logits, Y


def SSRN(hp, P, Y, training=False, gen=None, prefix="SSRN"):
    kw = dict(normtype=hp.norm, dropout_rate=hp.dropout_rate, training=training, gen=gen)
    i = 1
    t = conv1d(P, Y, "%s/C_%d" % (prefix, i), **kw); i += 1
    for j in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    for _ in range({4: 2, 8: 3}[hp.r]):
        t = conv1d_transpose(P, t, "%s/D_%d" % (prefix, i), hp.dropout_rate, training, gen); i += 1
        for j in range(2):
            t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 3 ** j, **kw); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), **kw); i += 1
    for _ in range(2):
        t = hc(P, t, "%s/HC_%d" % (prefix, i), 3, 1, **kw); i += 1
    t = conv1d(P, t, "%s/C_%d" % (prefix, i), **kw); i += 1
    for _ in range(2):
        t = conv1d(P, t, "%s/C_%d" % (prefix, i), act="relu", **kw); i += 1
    logits = conv1d(P, t, "%s/C_%d" % (prefix, i), **kw)
    Z = torch.sigmoid(logits) if getattr(hp, "squash_output_ssrn", True) else logits
    return logits, Z


def attention_guide(hp, dtype):
    n = torch.arange(hp.max_N, dtype=torch.float64)[:, None] / float(hp.max_N)
    t = torch.arange(hp.max_T, dtype=torch.float64)[None, :] / float(hp.max_T)
    return (1.0 - torch.exp(-(t - n) ** 2 / (2 * hp.g * hp.g))).to(torch.float32).to(dtype)


def text2mel_forward(hp, P, L, mels, mode="train", prev_max_attentions=None, K=None, V=None, gen=None):
    training = (mode == "train")
    S = torch.cat([torch.zeros_like(mels[:, :1]), mels[:, :-1]], 1)
    if K is None:
        K, V = TextEnc(hp, P, L, training, gen)
    Q = AudioEnc(hp, P, S, training, gen)
    R, ali, mx = Attention(hp, Q, K, V, mode == "synthesize", prev_max_attentions)
    logits, Y = AudioDec(hp, P, R, training, gen)
    return dict(K=K, V=V, Q=Q, R=R, alignments=ali, max_attentions=mx, Y_logits=logits, Y=Y)


def _lw(hp, which):
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
    """tf.pad(..., constant_values=value)[:, :max_N, :max_T] (architectures.py:261-263)."""
    return F.pad(x, (0, max_T, 0, max_N), value=float(value))[:, :max_N, :max_T]


def text2mel_loss(hp, out, mels, gts=None):
    """gts: per-utterance guides / forced-alignment targets of the batch [B, Ng, Tg] (architectures.py:57-58), else
    the global guide; hp.attention_guide_fa selects the MSE branch (:271-280)."""
    Y, logits, A = out["Y"], out["Y_logits"], out["alignments"]
    l2 = ((Y - mels) ** 2).mean()
    l1 = (Y - mels).abs().mean()
    bd = F.binary_cross_entropy_with_logits(logits, mels) if getattr(hp, "squash_output_t2m", True) else l1 * 0
    W = attention_guide(hp, Y.dtype)
    Nb, Tb = min(A.shape[1], hp.max_N), min(A.shape[2], hp.max_T)
    Ac = A[:, :Nb, :Tb]
    att = (Ac * W[None, :Nb, :Tb]).abs().sum() / float(Ac.numel())
    if gts is not None:
        gts = torch.as_tensor(gts).to(Y.dtype)
        mask = (_pad_crop(A, -1.0, hp.max_N, hp.max_T) != -1).to(Y.dtype)
        if getattr(hp, "attention_guide_fa", False):
            d = _pad_crop(A, 0.0, hp.max_N, hp.max_T) - _pad_crop(gts, 0.0, hp.max_N, hp.max_T)
            att = (d * d).sum() / mask.sum()
        else:
            att = ((_pad_crop(A, -1.0, hp.max_N, hp.max_T) * _pad_crop(gts, 1.0, hp.max_N, hp.max_T)).abs() * mask).sum() / mask.sum()
    w1, wbd, watt, w2 = _lw(hp, "t2m")
    loss = w1 * l1 + wbd * bd + watt * att + w2 * l2
    lw_cdp, lw_ain, lw_aout = (getattr(hp, n, 0.0) for n in ("lw_cdp", "lw_ain", "lw_aout"))
    if lw_cdp != 0.0 or lw_ain != 0.0 or lw_aout != 0.0:                        # architectures.py:283-321, 333-355
        Bn, Nn, Tn = A.shape
        s_in = A.sum(2)
        cdp = torch.log(1.0 + (1.0 - s_in) ** 2).sum() / (Bn * Nn)
        plogp = lambda P: torch.where(P != 0, P * torch.log(torch.where(P != 0, P, torch.ones_like(P))), torch.zeros_like(P))
        ain = -plogp(A / s_in[:, :, None]).sum() / (Bn * Nn) / math.log(Tn)
        aout = -plogp(A / A.sum(1, keepdim=True)).sum() / (Bn * Tn) / math.log(Nn)
        lw = getattr(hp, "loss_weights", None)
        if not (lw and "t2m" in lw):
            loss = loss + lw_cdp * cdp + lw_ain * ain + lw_aout * aout
        return [loss, l1, bd, att, l2, cdp, ain, aout]
    return [loss, l1, bd, att, l2]


def ssrn_loss(hp, logits, Z, mags):
    l2 = ((Z - mags) ** 2).mean()
    l1 = (Z - mags).abs().mean()
    bd = F.binary_cross_entropy_with_logits(logits, mags) if getattr(hp, "squash_output_ssrn", True) else l1 * 0
    w1, wbd, w2 = _lw(hp, "ssrn")
    return [w1 * l1 + wbd * bd + w2 * l2, l1, bd, l2]


def noam(init_lr, global_step, warmup=4000.0):
    step = float(global_step + 1)
    return init_lr * warmup ** 0.5 * min(step * warmup ** -1.5, step ** -0.5)


class TFAdam(object):
    """clip_by_value(+-1) then tf.train.AdamOptimizer (architectures.py:110-128)."""
    def __init__(self, hp, P):
        self.hp, self.P = hp, P
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}
        self.global_step = 0

    def step(self, grads, grad_scale=1.0):
        hp = self.hp
        lr = noam(hp.lr, self.global_step) if hp.decay_lr else hp.lr
        t = self.global_step + 1
        lr_t = lr * math.sqrt(1 - hp.beta2 ** t) / (1 - hp.beta1 ** t)
        with torch.no_grad():
            for k, p in self.P.items():
                g = grads.get(k)
                if g is None:
                    continue
                g = (g * grad_scale).clamp(-1.0, 1.0)
                self.m[k].mul_(hp.beta1).add_(g, alpha=1 - hp.beta1)
                self.v[k].mul_(hp.beta2).addcmul_(g, g, value=1 - hp.beta2)
                p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + hp.epsilon))
        self.global_step += 1


def text2mel_train_step(hp, P, opt, L, mels, gen=None, gts=None):
    """One `sess.run([global_step, loss_components, train_op])` (train.py:273)."""
    for p in P.values():
        p.grad = None
    out = text2mel_forward(hp, P, L, mels, "train", gen=gen)
    comps = text2mel_loss(hp, out, mels, gts=gts)
    comps[0].backward()
    grads = {k: p.grad for k, p in P.items() if p.grad is not None}
    opt.step(grads)
    return [float(c.detach()) for c in comps], grads


def ssrn_train_step(hp, P, opt, mels, mags, gen=None):
    for p in P.values():
        p.grad = None
    logits, Z = SSRN(hp, P, mels, True, gen)
    comps = ssrn_loss(hp, logits, Z, mags)
    comps[0].backward()
    grads = {k: p.grad for k, p in P.items() if p.grad is not None}
    opt.step(grads)
    return [float(c.detach()) for c in comps], grads


def synth_codedtext2mel(hp, P, K, V, ends, truncate=True, return_margin=False):
    """synthesize.py:150-230: the autoregressive loop that re-runs the whole graph per frame (`sess.run([g.Y,
    g.max_attentions, g.alignments], {g.K, g.V, g.mels, g.prev_max_attentions})`, :181-183), keeps row j of its outputs
    (:204-209), counts sentence ends (`endcount_threshold=1`, :163-165, :218-223) and stops when every sentence ended
    (:225-228).  Returns (Y [B,max_T,n_mels], t_ends, alignments [B,max_N,max_T]) like the reference.

    truncate=True feeds only rows 0..j of the mel buffer at step j.  The reference feeds all max_T rows but uses row j
    only; AudioEnc / AudioDec are causal, LayerNorm and the attention softmax act per row, and the window mask depends on
    N only (networks.py:304-313), so rows > j cannot influence row j: the truncated run is the same function (checked
    against truncate=False and against the numpy loop in tests/test_oracle.py) at half the cost, which is what lets the
    BASELINE synthesis shape (10 x N=150 x T=200) run in a test.
    return_margin=True also returns the smallest gap between the best and the second best score inside the attention
    window over all (sentence, frame) decisions that were used: a parity test can then tell a near-tie from a bug."""
    K, V = torch.as_tensor(K), torch.as_tensor(V)
    dt = K.dtype
    B, N = K.shape[0], K.shape[1]
    assert N == hp.max_N
    T = hp.max_T
    Y = torch.zeros(B, T, hp.n_mels, dtype=dt)
    alignments = torch.zeros(B, N, T, dtype=dt)
    prev = torch.zeros(B, dtype=torch.long)
    ends = np.asarray(ends)
    endcounts = np.zeros(ends.shape, dtype=int)
    t_ends = np.ones(ends.shape, dtype=int) * T
    margin = float("inf")
    n_idx = torch.arange(N)[None, :]
    with torch.no_grad():
        for j in range(T):
            rows = j + 1 if truncate else T
            mels = Y[:, :rows]
            S = torch.cat([torch.zeros_like(mels[:, :1]), mels[:, :-1]], 1)              # architectures.py:191
            Q = AudioEnc(hp, P, S)
            A = torch.matmul(Q, K.transpose(1, 2)) * (1.0 / math.sqrt(float(hp.d)))      # networks.py:300
            pv = prev.reshape(B, 1)
            key_masks = n_idx < pv                                                       # :304-306
            reverse = torch.flip(n_idx < (hp.max_N - hp.attention_win_size - pv), dims=[1])
            masks = (key_masks | reverse)[:, None, :].expand_as(A)
            A = torch.where(masks, torch.full_like(A, MASK_VALUE), A)
            if return_margin:
                top2 = torch.topk(A[:, j], 2, dim=-1).values
                margin = min(margin, float((top2[:, 0] - top2[:, 1]).min()))
            A = torch.softmax(A, -1)
            mx = A.argmax(-1)
            R = torch.matmul(A, V)
            if getattr(hp, "concatenate_query", True):
                R = torch.cat([R, Q], -1)
            _, Yj = AudioDec(hp, P, R)
            Y[:, j] = Yj[:, j]
            alignments[:, :, j] = A[:, j]
            prev = mx[:, j]
            endcounts += (prev.numpy() >= ends)
            for i in range(B):
                if t_ends[i] == T and endcounts[i] >= 1:
                    t_ends[i] = j
            if (t_ends < T).all():
                break
    out = (Y.numpy(), t_ends.tolist(), alignments.numpy())
    return out + (margin,) if return_margin else out


def advancing_keys(K, c=8.0, seed=0, period=8):
    """Synthetic "diagonal bias" for synthesis tests with random-init weights (SURVEY 8(d), C4 parity variant): with such
    weights the attention argmax never leaves the first window, so no sentence ever ends.  Adding c * (n mod period) * u
    (u a fixed random unit vector, n the key position) to the keys makes the best key inside the window
    [prev, prev + win) the last one whenever Q[t].u > 0 and the first one otherwise (the middle one where the window
    straddles a wrap of n mod period): the attention advances at a sentence- and frame-dependent pace and reaches the
    sentence ends in the middle of the run, while the scores stay of order c * period.  numpy [B, N, d] in and out."""
    K = np.asarray(K)
    rng = np.random.default_rng(seed)
    u = rng.standard_normal(K.shape[-1])
    u /= np.linalg.norm(u)
    n = (np.arange(K.shape[1]) % period).astype(np.float64)[None, :, None]
    return (K + c * n * u[None, None, :]).astype(K.dtype)
