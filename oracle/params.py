"""Variable inventory + initialisers of the reference graphs (TEST INFRASTRUCTURE ONLY).

Names/shapes follow what TF-1.12 creates at `modules.py:34-37,134-136,193,243-250` under the scopes
opened in `architectures.py:140,189-238` / `networks.py` (known answers: `train.py:194`:
'Text2Mel/TextEnc/embed_1/lookup_table' (V,128), 'Text2Mel/TextEnc/C_2/conv1d/kernel' (1,128,512)).
Initialisers: truncated_normal(0.1) for embeddings (`modules.py:37`),
variance_scaling_initializer() (factor 2, FAN_IN, truncated normal => sigma = sqrt(1.3*2/fan_in)) for kernels,
zeros for biases/beta, ones for gamma.
"""
import numpy as np


def _conv(out, prefix, name, k, cin, cout):
    s = "%s/%s" % (prefix, name)
    out.append((s + "/conv1d/kernel", (k, cin, cout), "kernel"))
    out.append((s + "/conv1d/bias", (cout,), "zeros"))
    out.append((s + "/normalize/beta", (cout,), "zeros"))
    out.append((s + "/normalize/gamma", (cout,), "ones"))


def _hc(out, prefix, name, k, c):
    s = "%s/%s" % (prefix, name)
    out.append((s + "/conv1d/kernel", (k, c, 2 * c), "kernel"))
    out.append((s + "/conv1d/bias", (2 * c,), "zeros"))
    for h in ("H1", "H2"):
        out.append((s + "/%s/beta" % h, (c,), "zeros"))
        out.append((s + "/%s/gamma" % h, (c,), "ones"))


def _deconv(out, prefix, name, c):
    s = "%s/%s" % (prefix, name)
    out.append((s + "/conv2d_transpose/kernel", (1, 3, c, c), "kernel_t"))
    out.append((s + "/conv2d_transpose/bias", (c,), "zeros"))
    out.append((s + "/normalize/beta", (c,), "zeros"))
    out.append((s + "/normalize/gamma", (c,), "ones"))


def text2mel_specs(hp):
    V, e, d, nm = len(hp.vocab), hp.e, hp.d, hp.n_mels
    out = []
    p = "Text2Mel/TextEnc"
    out.append((p + "/embed_1/lookup_table", (V, e), "embed"))
    _conv(out, p, "C_2", 1, e, 2 * d)
    _conv(out, p, "C_3", 1, 2 * d, 2 * d)
    for i in range(4, 14):
        _hc(out, p, "HC_%d" % i, 3, 2 * d)
    for i in range(14, 16):
        _hc(out, p, "HC_%d" % i, 1, 2 * d)
    p = "Text2Mel/AudioEnc"
    _conv(out, p, "C_1", 1, nm, d)
    _conv(out, p, "C_2", 1, d, d)
    _conv(out, p, "C_3", 1, d, d)
    for i in range(4, 14):
        _hc(out, p, "HC_%d" % i, 3, d)
    p = "Text2Mel/AudioDec"
    _conv(out, p, "C_1", 1, 2 * d, d)
    for i in range(2, 8):
        _hc(out, p, "HC_%d" % i, 3, d)
    for i in range(8, 11):
        _conv(out, p, "C_%d" % i, 1, d, d)
    _conv(out, p, "C_11", 1, d, nm)
    return out


def ssrn_specs(hp):
    c, nm, F = hp.c, hp.n_mels, hp.full_dim
    out = []
    p = "SSRN"
    i = 1
    _conv(out, p, "C_%d" % i, 1, nm, c); i += 1
    for _ in range(2):
        _hc(out, p, "HC_%d" % i, 3, c); i += 1
    for _ in range({4: 2, 8: 3}[hp.r]):
        _deconv(out, p, "D_%d" % i, c); i += 1
        for _ in range(2):
            _hc(out, p, "HC_%d" % i, 3, c); i += 1
    _conv(out, p, "C_%d" % i, 1, c, 2 * c); i += 1
    for _ in range(2):
        _hc(out, p, "HC_%d" % i, 3, 2 * c); i += 1
    _conv(out, p, "C_%d" % i, 1, 2 * c, F); i += 1
    for _ in range(3):
        _conv(out, p, "C_%d" % i, 1, F, F); i += 1
    return out


def _trunc_normal(rng, shape, std):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return x * std


def init_params(specs, seed=0, perturb=False):
    """perturb=True gives non-trivial bias/beta/gamma so LN/bias bugs are visible."""
    rng = np.random.default_rng(seed)
    P = {}
    for name, shape, kind in specs:
        if kind == "embed":
            v = _trunc_normal(rng, shape, 0.1)
        elif kind == "kernel":
            fan_in = shape[0] * shape[1]
            v = _trunc_normal(rng, shape, np.sqrt(1.3 * 2.0 / fan_in))
        elif kind == "kernel_t":          # [1,3,Cout,Cin]: conv2d_transpose fan_in = h*w*shape[2]
            fan_in = shape[0] * shape[1] * shape[2]
            v = _trunc_normal(rng, shape, np.sqrt(1.3 * 2.0 / fan_in))
        elif kind == "zeros":
            v = rng.uniform(-0.2, 0.2, shape) if perturb else np.zeros(shape)
        elif kind == "ones":
            v = rng.uniform(0.7, 1.3, shape) if perturb else np.ones(shape)
        else:
            raise ValueError(kind)
        P[name] = v.astype(np.float32)
    return P


class HP(object):
    """Minimal hyper-parameter bag with the `config/lj_test.cfg` hot-path values."""
    def __init__(self, **kw):
        self.vocab = ["<PADDING>"] + ["s%d" % i for i in range(64)]     # 65 symbols (lj_test.cfg:39-44)
        self.e, self.d, self.c = 128, 256, 512
        self.n_mels, self.full_dim, self.r = 80, 1025, 4
        self.norm = "layer"
        self.dropout_rate = 0.05
        self.max_N, self.max_T = 180, 210
        self.attention_win_size, self.g = 3, 0.2
        self.concatenate_query = True
        self.squash_output_t2m = self.squash_output_ssrn = True
        self.turn_off_monotonic_for_synthesis = False
        self.lw_mel = self.lw_bd1 = self.lw_att = 0.3333
        self.lw_t2m_l2 = 0.0
        self.lw_mag = self.lw_bd2 = 0.5
        self.lw_ssrn_l2 = 0.0
        self.lr, self.beta1, self.beta2, self.epsilon = 0.001, 0.9, 0.999, 1e-8
        self.decay_lr = True
        self.__dict__.update(kw)


def synthetic_batch(hp, B, N, T, seed=1234, text_len=None, with_mags=False, ragged=False):
    """SURVEY 8(d) synthetic inputs reproducing the data_load contract (`data_load.py:238-241,534-541`)."""
    rng = np.random.default_rng(seed)
    V = len(hp.vocab)
    L = np.zeros((B, N), np.int32)
    for b in range(B):
        n = text_len if text_len is not None else int(rng.integers(max(1, (2 * N) // 3), N + 1))
        L[b, :n] = rng.integers(1, V, n)
    mels = rng.uniform(1e-8, 1.0, (B, T, hp.n_mels)).astype(np.float32)
    if ragged:
        for b in range(B):
            tl = int(rng.integers(T // 2, T + 1))
            mels[b, tl:] = 0
    out = dict(L=L, mels=mels)
    if with_mags:
        out["mags"] = rng.uniform(1e-8, 1.0, (B, T * hp.r, hp.full_dim)).astype(np.float32)
    return out
