"""Layer primitives with the reference's names and signatures (`modules.py:15-258` of CSTR-Edinburgh/ophelia):
embed, normalize, conv1d, hc, conv1d_transpose.  Tensors are fp32 CUDA `[B, time, C]`; all arithmetic runs in
libophelia_sm100.so (fused tcgen05 implicit-GEMM + row-wise kernels).  There is no CPU path.

Differences from the TF graph-mode originals are keyword-only extras used for fusion (`out=`, `in_shift=`,
`want_sigmoid=`): positional/keyword use as in the reference's networks.py works unchanged.
"""
import torch

from . import ops
from .variables import get_store, layer_seed, scoped, variable_scope

relu = "relu"          # stands in for tf.nn.relu at the call sites of networks.py


# ------------------------------------------------------------------------------------------------ autodiff tape
class Tape(object):
    """Reverse-mode tape for one chain of layers (TextEnc, AudioEnc, AudioDec and SSRN are pure chains)."""
    current = None

    def __init__(self):
        self.steps = []

    def __enter__(self):
        self._prev, Tape.current = Tape.current, self
        return self

    def __exit__(self, *exc):
        Tape.current = self._prev

    def backward(self, grad, release=True, marks=None):
        """release=False keeps the closures (and the activations they hold) alive: needed while weight-gradient
        GEMMs on a side stream may still read them; call release() after joining that stream.
        marks = {n: callback}: callback() runs once the backward functions of the LAST n layers are enqueued (gradient
        buckets of those layers can then be all-reduced while the earlier layers' backward still runs)."""
        for k, fn in enumerate(reversed(self.steps), 1):
            grad = fn(grad)
            if marks and k in marks:
                marks[k]()
        if release:
            self.steps = []
        return grad

    def release(self):
        self.steps = []


def _record(fn):
    if Tape.current is not None:
        Tape.current.steps.append(fn)


def _recording(training):
    return bool(training) and Tape.current is not None


def _packed(store, name, deconv=False):
    """Lazily (re)built split-bf16 image of a conv kernel; re-packed when the store's values changed."""
    pk = store.packed.get(name)
    if pk is None:
        pk = ops.PackedConv(store.get(name), deconv=deconv)
        pk.version = store.version
        store.packed[name] = pk
    elif pk.version != store.version:
        pk.repack()
        pk.version = store.version
    return pk


def _step_ptr(store, training):
    return getattr(store, "global_step", None) if training else None


def _padding_code(padding):
    p = padding.lower()
    if p == "causal":
        return ops.CAUSAL
    if p == "same":
        return ops.SAME
    raise ValueError("padding %r is not used on the dc_tts path" % padding)


# ------------------------------------------------------------------------------------------------ embed
def embed(inputs, vocab_size, num_units, zero_pad=True, scope="embedding", reuse=None):
    """modules.py:15-44.  inputs int32 [B, N] -> [B, N, num_units]; row 0 of the table reads as zeros."""
    assert zero_pad, "the path always uses zero_pad=True"
    store = get_store()
    with variable_scope(scope, reuse=reuse):
        name = scoped("lookup_table")
        store.declare(name, (vocab_size, num_units), "embed")
    store.finalize()
    ids = inputs.to(torch.int32).contiguous()
    out = ops.embed_fwd(ids, store.get(name))
    if Tape.current is not None:
        def bwd(dout):
            ops.embed_bwd(ids, dout, store.grad(name))
            return None
        _record(bwd)
    return out


def normalize(inputs, scope="normalize", reuse=None, normtype='layer'):
    """modules.py:47-75: layer norm over the last axis (tf.contrib.layers.layer_norm: biased variance, eps 1e-12, gamma /
    beta of shape [C] under `scope`), or the identity for normtype None.  Inside conv1d / hc / conv1d_transpose the same
    arithmetic runs fused into the layer's tail; this is the stand-alone function (one launch)."""
    assert normtype in (None, 'layer'), "batch norm is unused by every shipped config"
    if normtype is None:
        return inputs
    store = get_store()
    C = inputs.shape[-1]
    with variable_scope(scope, reuse=reuse):
        gn, ben = scoped("gamma"), scoped("beta")
        store.declare(ben, (C,), "zeros")
        store.declare(gn, (C,), "ones")
    store.finalize()
    gamma, beta = store.get(gn), store.get(ben)
    rec = Tape.current is not None
    y, stats = ops.normalize_fwd(inputs, gamma, beta, save=rec)
    if rec:
        def bwd(dy):
            return ops.normalize_bwd(dy, inputs, stats, gamma, beta, store.grad(gn), store.grad(ben))
        _record(bwd)
    return y


# ------------------------------------------------------------------------------------------------ conv1d
def conv1d(inputs, filters=None, size=1, rate=1, padding="SAME", dropout_rate=0, use_bias=True, activation_fn=None,
           training=True, scope="conv1d", reuse=None, normtype='layer', lcc=0, codes=None,
           *, out=None, in_shift=0, want_sigmoid=False, planes=True):
    """modules.py:91-146: conv (+bias) -> layer norm -> activation -> dropout (training only) -> optional per-speaker
    channel gates (lcc = number of speaker codes, codes int [B, 1]; modules.py:78-88, 141-144)."""
    assert use_bias, "the bias-free variant is unused by the path"
    assert normtype in (None, 'layer'), "batch norm is unused by every shipped config"
    assert activation_fn in (None, relu)
    store = get_store()
    cin = inputs.shape[-1]
    if filters is None:
        filters = cin
    with variable_scope(scope):
        kn, bn = scoped("conv1d/kernel"), scoped("conv1d/bias")
        gn, ben = scoped("normalize/gamma"), scoped("normalize/beta")
        store.declare(kn, (size, cin, filters), "kernel")
        store.declare(bn, (filters,), "zeros")
        if normtype == 'layer':
            store.declare(ben, (filters,), "zeros")
            store.declare(gn, (filters,), "ones")
        ln_ = scoped("lcc_embed/lookup_table")
        if lcc:
            store.declare(ln_, (lcc, filters), "embed")
    store.finalize()
    norm = normtype == 'layer'
    pk = _packed(store, kn)
    gamma = store.get(gn) if norm else None
    beta = store.get(ben) if norm else None
    act = ops.ACT_RELU if activation_fn == relu else ops.ACT_NONE
    pad = _padding_code(padding)
    drop = float(dropout_rate) if training else 0.0
    seed = layer_seed(kn, store)
    step = _step_ptr(store, training)
    rec = _recording(training)
    gated = bool(lcc)
    y, ysig, saved = ops.conv1d_fwd(inputs, pk, store.get(bn), gamma, beta, rate, pad, in_shift, act, norm, drop, seed,
                                    step, save=rec, y=None if gated else out, want_sigmoid=want_sigmoid and not gated,
                                    planes=planes and out is None and not gated)
    y0 = y
    if gated:
        assert out is None and codes is not None
        cds = codes.to(torch.int32).reshape(-1).contiguous()
        y, ysig = ops.lcc_fwd(y0, store.get(ln_), cds, want_sigmoid=want_sigmoid, planes=planes)
    if rec:
        need_dx = not getattr(inputs, "_oph_no_grad", False)

        def bwd(dy):
            if gated:
                dy = ops.lcc_bwd(dy, y0, store.get(ln_), cds, store.grad(ln_))
            return ops.conv1d_bwd(dy, inputs, saved, _packed(store, kn), gamma, beta, store.grad(kn), store.grad(bn),
                                  store.grad(gn) if norm else None, store.grad(ben) if norm else None, rate, pad,
                                  in_shift, act, norm, drop, seed, step, need_dx=need_dx)
        _record(bwd)
    if want_sigmoid:
        return y, ysig
    return y


# ------------------------------------------------------------------------------------------------ highway conv
def hc(inputs, filters=None, size=1, rate=1, padding="SAME", dropout_rate=0, use_bias=True, activation_fn=None,
       training=True, scope="hc", reuse=None, normtype='layer', lcc=0, codes=None, *, out=None, out_planes=None):
    """modules.py:148-207: conv to 2C, split, LN(H1), LN(H2), gate = sigmoid(H1), gate*H2 + (1-gate)*inputs."""
    assert use_bias and activation_fn is None
    assert normtype in (None, 'layer')
    store = get_store()
    C = inputs.shape[-1]
    if filters is None:
        filters = C
    assert filters == C, "highway conv keeps the channel count"
    with variable_scope(scope):
        kn, bn = scoped("conv1d/kernel"), scoped("conv1d/bias")
        names = [scoped("H1/gamma"), scoped("H1/beta"), scoped("H2/gamma"), scoped("H2/beta")]
        store.declare(kn, (size, C, 2 * C), "kernel")
        store.declare(bn, (2 * C,), "zeros")
        if normtype == 'layer':
            store.declare(names[1], (C,), "zeros")
            store.declare(names[0], (C,), "ones")
            store.declare(names[3], (C,), "zeros")
            store.declare(names[2], (C,), "ones")
        ln_ = scoped("lcc_embed/lookup_table")
        if lcc:
            store.declare(ln_, (lcc, C), "embed")
    store.finalize()
    gate = None
    if lcc:                 # per-speaker gates on the transformation connection (modules.py:200-201)
        assert codes is not None
        gate = (store.get(ln_), codes.to(torch.int32).reshape(-1).contiguous())
    norm = normtype == 'layer'
    pk = _packed(store, kn)
    g1, b1, g2, b2 = ([store.get(n) for n in names] if norm else [None] * 4)
    pad = _padding_code(padding)
    drop = float(dropout_rate) if training else 0.0
    seed = layer_seed(kn, store)
    step = _step_ptr(store, training)
    rec = _recording(training)
    y, saved = ops.hc_fwd(inputs, pk, store.get(bn), g1, b1, g2, b2, rate, pad, norm, drop, seed, step, save=rec, y=out,
                          planes=True, y_planes=out_planes, lcc=gate)
    if rec:
        def bwd(dy):
            gr = [store.grad(n) for n in names] if norm else [None] * 4
            return ops.hc_bwd(dy, inputs, saved, _packed(store, kn), g1, b1, g2, b2, store.grad(kn), store.grad(bn),
                              gr[0], gr[1], gr[2], gr[3], rate, pad, norm, drop, seed, step, lcc=gate,
                              dtable=store.grad(ln_) if gate else None)
        _record(bwd)
    return y


# ------------------------------------------------------------------------------------------------ transposed conv
def conv1d_transpose(inputs, filters=None, size=3, stride=2, padding='same', dropout_rate=0, use_bias=True,
                     activation=None, training=True, scope="conv1d_transpose", reuse=None, normtype='layer'):
    """modules.py:209-258: time x2 up-sampling; layer norm is always on (the caller never passes hp.norm)."""
    assert size == 3 and stride == 2 and padding.lower() == 'same' and use_bias and activation is None
    assert normtype == 'layer'
    store = get_store()
    C = inputs.shape[-1]
    if filters is None:
        filters = C
    assert filters == C
    with variable_scope(scope, reuse=reuse):
        kn, bn = scoped("conv2d_transpose/kernel"), scoped("conv2d_transpose/bias")
        gn, ben = scoped("normalize/gamma"), scoped("normalize/beta")
        store.declare(kn, (1, 3, C, C), "kernel_t")
        store.declare(bn, (C,), "zeros")
        store.declare(ben, (C,), "zeros")
        store.declare(gn, (C,), "ones")
    store.finalize()
    pk = _packed(store, kn, deconv=True)
    gamma, beta = store.get(gn), store.get(ben)
    drop = float(dropout_rate) if training else 0.0
    seed = layer_seed(kn, store)
    step = _step_ptr(store, training)
    rec = _recording(training)
    y, saved = ops.deconv_fwd(inputs, pk, store.get(bn), gamma, beta, drop, seed, step, save=rec)
    if rec:
        def bwd(dy):
            return ops.deconv_bwd(dy, inputs, saved, _packed(store, kn, deconv=True), gamma, beta, store.grad(kn),
                                  store.grad(bn), store.grad(gn), store.grad(ben), drop, seed, step)
        _record(bwd)
    return y
