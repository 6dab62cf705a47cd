"""The exact incremental frame step, restated in numpy with the kernels' own index formulas, against the oracle's full
re-computation loop (synthesize.py:150-230).

What is exact (hard part 5): AudioEnc is causal and its input row t is final once frame t-1 exists, so Q[t]
can be cached -- every layer keeps its output history and a step computes row j only (csrc/arstep.cuh: tap i reads row
j - in_shift - (k - 1 - i) * rate, the reduction over k * Cin is cut into slices).  Attention is NOT cacheable: the
window mask derived from the latest prev_max_attentions is applied to every time row (networks.py:304-313), so the context
rows R[t < j] change whenever the window moves, and AudioDec's row j sees them through its 84-frame causal reach.  Those
two are therefore re-run per step, but only over rows [max(0, j - reach), j].  The naive variant that also caches R and
the AudioDec states diverges from the reference as soon as the window moves (second half of the test)."""
import numpy as np

from helpers import make_hp, maxabs, oracle_params
from oracle import dctts_numpy as on
from oracle.params import synthetic_batch


def _ln(z, gamma, beta):
    mean = z.mean(-1, keepdims=True)
    var = ((z - mean) ** 2).mean(-1, keepdims=True)
    return (z - mean) / np.sqrt(var + 1e-12) * gamma + beta


def _gemv(x_hist, W, j, in_shift, rate, kslice):
    """ar_gemv_kernel: partial sums over slices of the (tap, channel) reduction, summed in slice order by the tail."""
    k, Cin, O = W.shape
    B = x_hist.shape[0]
    K = k * Cin
    Wf = W.reshape(K, O)
    xs = np.zeros((B, K))
    for kk in range(K):
        tap, c = kk // Cin, kk % Cin
        t = j - in_shift - (k - 1 - tap) * rate
        if t >= 0:
            xs[:, kk] = x_hist[:, t, c]
    z = np.zeros((B, O))
    for k0 in range(0, K, kslice):
        z += xs[:, k0:k0 + kslice] @ Wf[k0:k0 + kslice]
    return z


def _layer_row(P, prefix, spec, x, y, j, in_shift=0):
    """ar_tail_kernel after ar_gemv_kernel: row j of one layer from the history x of the layer below."""
    scope, kind, k, rate, act = spec
    p = "Text2Mel/%s/%s" % (prefix, scope)
    z = _gemv(x, P[p + "/conv1d/kernel"], j, in_shift, rate, 48) + P[p + "/conv1d/bias"]
    if kind == "hc":
        C = x.shape[2]
        h1 = _ln(z[:, :C], P[p + "/H1/gamma"], P[p + "/H1/beta"])
        h2 = _ln(z[:, C:], P[p + "/H2/gamma"], P[p + "/H2/beta"])
        g = on.sigmoid(h1)
        y[:, j] = g * h2 + (1 - g) * x[:, j]
    else:
        u = _ln(z, P[p + "/normalize/gamma"], P[p + "/normalize/beta"])
        y[:, j] = np.maximum(u, 0) if act else u
    return y


def incremental_loop(hp, P, K, V, ends, synth, cache_decoder=False):
    B, N, d = K.shape
    T, nm = hp.max_T, hp.n_mels
    enc, dec = synth._frame_step_layers(hp)
    reach = synth.decoder_reach(hp)
    W = min(T, reach + 1)
    Y = np.zeros((B, T, nm)); ali = np.zeros((B, N, T)); prev = np.zeros(B, int)
    h_enc = [np.zeros((B, T, d)) for _ in enc]
    h_dec = [np.zeros((B, T, d)) for _ in dec[:-1]] + [np.zeros((B, T, nm))]
    rq = np.zeros((B, T, 2 * d))
    ends = np.asarray(ends); endcounts = np.zeros(ends.shape, int); t_ends = np.ones(ends.shape, int) * T
    hp_w = synth._window_hp(hp, W)
    for j in range(T):
        x = Y
        for i, spec in enumerate(enc):
            x = _layer_row(P, "AudioEnc", spec, x, h_enc[i], j, in_shift=1 if i == 0 else 0)
        Q = x
        if cache_decoder:       # the tempting O(1) variant: row j of the attention and of every decoder layer only
            Rj, aj, mj = on.Attention(synth._window_hp(hp, 1), Q[:, j:j + 1], K, V, True, prev)
            rq[:, j] = Rj[:, 0]
            x = rq
            for i, spec in enumerate(dec):
                x = _layer_row(P, "AudioDec", spec, x, h_dec[i], j)
            Y[:, j] = on.sigmoid(x[:, j]); ali[:, :, j] = aj[:, :, 0]; prev = mj[:, 0]
        else:
            s = max(0, min(j - reach, T - W))                                # ar_window_gather_kernel
            r = j - s
            Rw, aw, mw = on.Attention(hp_w, Q[:, s:s + W], K, V, True, prev)
            _, Yw = on.AudioDec(hp, P, Rw)
            Y[:, j] = Yw[:, r]; ali[:, :, j] = aw[:, :, r]; prev = mw[:, r]  # ar_window_scatter_kernel
        endcounts += (prev >= ends)
        for i in range(B):
            if t_ends[i] == T and endcounts[i] >= 1:
                t_ends[i] = j
        if (t_ends < T).all():
            break
    return Y, t_ends.tolist(), ali


def test_incremental_frame_step_equals_full_recomputation():
    from ophelia_b200 import synthesize as synth
    hp = make_hp(max_N=12, max_T=110, d=32, e=16, n_mels=8)     # 110 frames: beyond the decoder's 84-frame reach
    assert synth.decoder_reach(hp) == 84
    P = oracle_params(hp, "t2m", seed=9)
    P = {k: np.asarray(v, np.float64) for k, v in P.items()}
    b = synthetic_batch(hp, 2, 12, 110, text_len=9)
    K, V = on.TextEnc(hp, P, b["L"])
    ends = np.array([99, 99])                                   # never reached: all 110 frames are generated
    Yr, tr, ar = on.synth_codedtext2mel(hp, P, K, V, ends)
    assert len(set(ar[0].argmax(0))) > 3                        # the window does move in this run
    Yi, ti, ai = incremental_loop(hp, P, K, V, ends, synth)
    assert ti == tr and maxabs(Yi, Yr) < 1e-9 and maxabs(ai, ar) < 1e-9
    # early stopping bookkeeping is the reference's
    ends = np.array([9, 9])
    Yr, tr, ar = on.synth_codedtext2mel(hp, P, K, V, ends)
    Yi, ti, ai = incremental_loop(hp, P, K, V, ends, synth)
    assert ti == tr and max(tr) < 109 and maxabs(Yi[:, :max(tr) + 1], Yr[:, :max(tr) + 1]) < 1e-9
    # caching the context rows and the decoder states as well is NOT the reference's function
    Yn, tn, an = incremental_loop(hp, P, K, V, np.array([99, 99]), synth, cache_decoder=True)
    Yr, tr, ar = on.synth_codedtext2mel(hp, P, K, V, np.array([99, 99]))
    assert maxabs(Yn[:, 0], Yr[:, 0]) < 1e-9 and maxabs(Yn, Yr) > 1e-2


def test_cluster_moment_combination_equals_two_pass_layer_norm():
    """ar_encoder_kernel: each of the 8 CTAs of a cluster owns C / 8 channels and contributes (mean, sum of squared
    deviations) of its slice; mean = average of the slice means, M2 = sum M2_r + (C / 8) sum (mean_r - mean)^2 (Chan et
    al.).  Same moments as the two global passes of tf.nn.moments, including for badly centred rows."""
    rng = np.random.default_rng(0)
    for C, offset in ((256, 0.0), (512, 300.0), (256, -1e4)):
        z = rng.normal(size=(5, C)) * 3.0 + offset
        parts = z.reshape(5, 8, C // 8)
        m_r = parts.mean(-1)
        M2_r = ((parts - m_r[..., None]) ** 2).sum(-1)
        mean = m_r.mean(-1)
        M2 = M2_r.sum(-1) + (C // 8) * ((m_r - mean[:, None]) ** 2).sum(-1)
        np.testing.assert_allclose(mean, z.mean(-1), rtol=1e-12)
        np.testing.assert_allclose(M2 / C, z.var(-1), rtol=1e-9)
