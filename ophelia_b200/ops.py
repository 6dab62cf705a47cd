"""Tensor-level wrappers over the C ABI (include/ophelia_b200.h).

torch is used for device memory and streams only; every arithmetic op below is one call into
libophelia_sm100.so.  Activations are fp32 CUDA tensors `[B, time, C]` whose rows may be strided
(`stride(-1) == 1`, `stride(0) == time * stride(1)`), which is how K/V, R/Q share buffers.
"""
import contextlib
import ctypes
import gc
import os

import torch

from . import _lib

SAME, CAUSAL = 0, 1
ACT_NONE, ACT_RELU = 0, 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


@contextlib.contextmanager
def capture(graph):
    """`with torch.cuda.graph(graph)` that cannot be invalidated by Python's cyclic collector.  A dead reference cycle
    may own device resources whose release is illegal while any capture is open (another captured step with its private
    pool, e.g. the per-shape autoregressive step a discarded Graph object kept): collected in the middle of a capture
    it turns every later launch into cudaErrorStreamCaptureInvalidated.  torch.cuda.graph no longer collects on entry,
    so collect here and keep the collector off until the capture has ended."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        prio = os.environ.get("OPH_CAPTURE_PRIORITY")     # diagnostics: priority of the capturing (main-chain) stream
        if prio is not None:
            with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=int(prio))):
                yield graph
        else:
            with torch.cuda.graph(graph):
                yield graph
    finally:
        if was_enabled:
            gc.enable()


# Weight-gradient GEMMs on a side stream (oph_wgrad_stream): buffers they read (dz) are parked in `_keepalive` until
# the caller has joined the side stream, so that the caching allocator cannot hand them to a later layer.
_keepalive = None


def set_wgrad_stream(stream):
    """stream: torch.cuda.Stream or None (in-order).  Returns nothing; per host thread."""
    global _keepalive
    if stream is None:
        _lib.call("oph_wgrad_stream", None, 0)
        _keepalive = None
    else:
        _lib.call("oph_wgrad_stream", stream.cuda_stream, 1)
        if _keepalive is None:
            _keepalive = []


def _park(*tensors):
    if _keepalive is not None:
        _keepalive.extend(t for t in tensors if t is not None)


def take_keepalive():
    """Hand the parked buffers to the caller (who drops them after joining the side stream)."""
    global _keepalive
    out, _keepalive = _keepalive, ([] if _keepalive is not None else None)
    return out


def _p(t):
    return None if t is None else t.data_ptr()


def _rows(t):
    """(ld, B, L, C) of a [B, L, C] activation; checks the layout contract."""
    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 3, "expected fp32 CUDA [B, time, C]"
    B, L, C = t.shape
    assert t.stride(2) == 1 and (B == 1 or t.stride(0) == L * t.stride(1)), "rows must be uniformly strided"
    ld = t.stride(1)
    assert ld % 4 == 0 and t.data_ptr() % 16 == 0, "rows must be 16-byte aligned"
    return ld, B, L, C


def _pad4(c):
    return (c + 3) // 4 * 4


def _act(t, use_planes=True, planes=None):
    """oph_act for a [B, L, C] activation; planes ride along as `t._oph_planes = (hi, lo)` (bf16 [B, L, C]) or are
    passed explicitly."""
    a = _lib.Act()
    ld = _rows(t)[0]
    a.f32, a.ld = t.data_ptr(), ld
    pl = planes if planes is not None else (getattr(t, "_oph_planes", None) if use_planes else None)
    if pl is not None:
        a.hi, a.lo, a.ldp = pl[0].data_ptr(), pl[1].data_ptr(), pl[0].stride(1)
    return a


def _out_act(y, want_planes, given=None):
    """oph_act for an output; allocates (and attaches) the planes when the width supports them, or uses `given`
    (hi, lo) views (e.g. one half of a wider planes buffer)."""
    a = _lib.Act()
    a.f32, a.ld = y.data_ptr(), y.stride(1)
    if given is not None:
        hi, lo = given
        a.hi, a.lo, a.ldp = hi.data_ptr(), lo.data_ptr(), hi.stride(1)
        y._oph_planes = (hi, lo)
    elif want_planes:                   # any width: plane rows are padded to 8 elements (16 bytes) for the tensor maps
        B, L, C = y.shape
        hi = torch.empty(B, L, _pad8(C), device=y.device, dtype=torch.bfloat16)[:, :, :C]
        lo = torch.empty(B, L, _pad8(C), device=y.device, dtype=torch.bfloat16)[:, :, :C]
        a.hi, a.lo, a.ldp = hi.data_ptr(), lo.data_ptr(), hi.stride(1)
        y._oph_planes = (hi, lo)
    return a


def _pad8(c):
    return (c + 7) // 8 * 8


def new_act(B, L, C, device):
    """[B, L, C] view over a buffer whose row stride is padded to a multiple of 8 floats (fp32 rows stay 16-byte
    aligned, and the same bytes can hold the two bf16 planes of a gradient with 16-byte aligned rows)."""
    ld = _pad8(C)
    if ld == C:
        return torch.empty(B, L, C, device=device, dtype=torch.float32)
    return torch.zeros(B, L, ld, device=device, dtype=torch.float32)[:, :, :C]


class PackedConv(object):
    """Split-bf16 shared-memory images of one conv kernel (forward GEMM + input-gradient GEMM)."""

    def __init__(self, w, deconv=False, need_bwd=True):
        self.deconv = bool(deconv)
        if deconv:
            assert w.dim() == 4 and w.shape[0] == 1 and w.shape[1] == 3      # [1,3,Cout,Cin]
            self.k, self.cout, self.cin = 3, int(w.shape[2]), int(w.shape[3])
        else:
            assert w.dim() == 3                                              # [k,Cin,Cout]
            self.k, self.cin, self.cout = (int(s) for s in w.shape)
        self.w = w
        dev = w.device
        self.fwd = torch.empty(_lib.pack_bytes(self.k, self.cin, self.cout, deconv, 0), dtype=torch.uint8, device=dev)
        self.bwd = (torch.empty(_lib.pack_bytes(self.k, self.cin, self.cout, deconv, 1), dtype=torch.uint8, device=dev)
                    if need_bwd else None)
        self.repack()

    def repack(self):
        assert self.w.is_contiguous()
        _lib.call("oph_conv_pack", _p(self.w), self.k, self.cin, self.cout, int(self.deconv), _p(self.fwd),
                  _p(self.bwd), _stream())


# ---------------------------------------------------------------------------------------------- conv1d
def conv1d_fwd(x, pk, bias, gamma, beta, rate=1, padding=SAME, in_shift=0, act=ACT_NONE, norm=True,
               drop_p=0.0, seed=0, step=None, save=False, y=None, want_sigmoid=False, planes=True):
    ldx, B, L, cin = _rows(x)
    assert cin == pk.cin
    cout = pk.cout
    dev = x.device
    z = new_act(B, L, cout, dev)
    stats = torch.empty(B * L, 2, device=dev, dtype=torch.float32) if (save and norm) else None
    if y is None:
        y = new_act(B, L, cout, dev)
    ysig = new_act(B, L, cout, dev) if want_sigmoid else None
    ya = _out_act(y, planes)
    xpl = getattr(x, "_oph_planes", None)
    if xpl is None:                        # inputs from outside the row-wise kernels (mels, embeddings): split once per call
        xpl = split_planes(x)
    _lib.call("oph_conv1d_fwd", _act(x, planes=xpl), _p(pk.fwd), _p(bias), _p(gamma), _p(beta), _p(z), z.stride(1), _p(stats),
              ya, _p(ysig), ysig.stride(1) if ysig is not None else 0, B, L, cin, cout, pk.k, rate,
              padding, in_shift, act, int(bool(norm)), float(drop_p), int(seed), _p(step), _stream())
    return y, ysig, (z, stats, xpl)


def conv1d_bwd(dy, x, saved, pk, gamma, beta, dw, dbias, dgamma, dbeta, rate=1, padding=SAME, in_shift=0,
               act=ACT_NONE, norm=True, drop_p=0.0, seed=0, step=None, need_dx=True, dx=None):
    z, stats, xpl = saved
    ldx, B, L, cin = _rows(x)
    lddy = _rows(dy)[0]
    dev = x.device
    dz = new_act(B, L, pk.cout, dev)
    if need_dx and dx is None:
        dx = new_act(B, L, cin, dev)
    _lib.call("oph_conv1d_bwd", _p(dy), lddy, _act(x, planes=xpl), _p(z), z.stride(1), _p(stats), _p(pk.bwd), _p(gamma),
              _p(beta), _p(dz), dz.stride(1), _p(dx) if need_dx else None, dx.stride(1) if need_dx else 0, _p(dw),
              _p(dbias), _p(dgamma), _p(dbeta), B, L, cin, pk.cout, pk.k, rate, padding, in_shift, act,
              int(bool(norm)), float(drop_p), int(seed), _p(step), _stream())
    _park(dz)
    return dx if need_dx else None


# ---------------------------------------------------------------------------------------------- normalize
def normalize_fwd(x, gamma, beta, save=False, planes=False):
    ldx, B, L, C = _rows(x)
    y = new_act(B, L, C, x.device)
    stats = torch.empty(B * L, 2, device=x.device, dtype=torch.float32) if save else None
    _lib.call("oph_normalize_fwd", _p(x), ldx, _p(gamma), _p(beta), _out_act(y, planes), _p(stats), B * L, C, _stream())
    return y, stats


def normalize_bwd(dy, x, stats, gamma, beta, dgamma, dbeta):
    ldx, B, L, C = _rows(x)
    lddy = _rows(dy)[0]
    dx = new_act(B, L, C, x.device)
    _lib.call("oph_normalize_bwd", _p(dy), lddy, _p(x), ldx, _p(stats), _p(gamma), _p(beta), _p(dx), dx.stride(1),
              _p(dgamma), _p(dbeta), B * L, C, _stream())
    return dx


# ---------------------------------------------------------------------------------------------- highway conv
def hc_fwd(x, pk, bias, g1, b1, g2, b2, rate=1, padding=SAME, norm=True, drop_p=0.0, seed=0, step=None,
           save=False, y=None, planes=True, y_planes=None, lcc=None):
    """lcc = (table [ncodes, C], codes int32 [B]): per-speaker gates on LN(H2) (modules.py:200-201)."""
    ldx, B, L, C = _rows(x)
    if lcc is not None:
        _lib.call("oph_lcc_context", _p(lcc[0]), _p(lcc[1]), None)
    assert pk.cin == C and pk.cout == 2 * C
    dev = x.device
    z = torch.empty(B, L, 2 * C, device=dev, dtype=torch.float32)
    stats = torch.empty(B * L, 4, device=dev, dtype=torch.float32) if (save and norm) else None
    if y is None:
        y = torch.empty(B, L, C, device=dev, dtype=torch.float32)
    ya = _out_act(y, planes, y_planes)
    _lib.call("oph_hc_fwd", _act(x), _p(pk.fwd), _p(bias), _p(g1), _p(b1), _p(g2), _p(b2), _p(z), z.stride(1),
              _p(stats), ya, B, L, C, pk.k, rate, padding, int(bool(norm)), float(drop_p),
              int(seed), _p(step), _stream())
    return y, (z, stats)


def hc_bwd(dy, x, saved, pk, g1, b1, g2, b2, dw, dbias, dg1, db1, dg2, db2, rate=1, padding=SAME, norm=True,
           drop_p=0.0, seed=0, step=None, dx=None, lcc=None, dtable=None):
    z, stats = saved
    ldx, B, L, C = _rows(x)
    scratch = None
    if lcc is not None:
        scratch = torch.empty(B * L, C, device=x.device, dtype=torch.float32)
        _lib.call("oph_lcc_context", _p(lcc[0]), _p(lcc[1]), _p(scratch))
    lddy = _rows(dy)[0]
    dev = x.device
    dz = torch.empty(B, L, 2 * C, device=dev, dtype=torch.float32)
    if dx is None:
        dx = torch.empty(B, L, C, device=dev, dtype=torch.float32)
    # (dxres / ldxr of the C ABI are unused: the residual-path gradient is written straight into dx)
    _lib.call("oph_hc_bwd", _p(dy), lddy, _act(x), _p(z), z.stride(1), _p(stats), _p(pk.bwd), _p(g1), _p(b1),
              _p(g2), _p(b2), _p(dz), dz.stride(1), None, 0, _p(dx), dx.stride(1), _p(dw),
              _p(dbias), _p(dg1), _p(db1), _p(dg2), _p(db2), B, L, C, pk.k, rate, padding, int(bool(norm)),
              float(drop_p), int(seed), _p(step), _stream())
    if lcc is not None:
        _lib.call("oph_lcc_reduce", _p(scratch), _p(lcc[1]), _p(dtable), B, L, C, _stream())
    _park(dz)
    return dx


# ---------------------------------------------------------------------------------------------- channel gates
def lcc_fwd(y0, table, codes, want_sigmoid=False, planes=True):
    """modules.learn_channel_contributions behind a conv1d layer: out = sigmoid(embed(codes)) * y0."""
    ld0, B, L, C = _rows(y0)
    out = new_act(B, L, C, y0.device)
    sig = new_act(B, L, C, y0.device) if want_sigmoid else None
    _lib.call("oph_lcc_fwd", _p(y0), ld0, _p(table), _p(codes), _out_act(out, planes), _p(sig),
              sig.stride(1) if sig is not None else 0, B, L, C, _stream())
    return out, sig


def lcc_bwd(dy, y0, table, codes, dtable):
    ld0, B, L, C = _rows(y0)
    lddy = _rows(dy)[0]
    dy0 = new_act(B, L, C, y0.device)
    scratch = torch.empty(B * L, C, device=y0.device, dtype=torch.float32)
    _lib.call("oph_lcc_bwd", _p(dy), lddy, _p(y0), ld0, _p(table), _p(codes), _p(dy0), dy0.stride(1), _p(scratch), B, L, C,
              _stream())
    _lib.call("oph_lcc_reduce", _p(scratch), _p(codes), _p(dtable), B, L, C, _stream())
    return dy0


# ---------------------------------------------------------------------------------------------- transposed conv
def deconv_fwd(x, pk, bias, gamma, beta, drop_p=0.0, seed=0, step=None, save=False):
    ldx, B, L, C = _rows(x)
    assert pk.deconv and pk.cin == C and pk.cout == C
    dev = x.device
    z = torch.empty(B, 2 * L, C, device=dev, dtype=torch.float32)
    stats = torch.empty(B * 2 * L, 2, device=dev, dtype=torch.float32) if save else None
    y = torch.empty(B, 2 * L, C, device=dev, dtype=torch.float32)
    ya = _out_act(y, True)
    _lib.call("oph_deconv_fwd", _act(x), _p(pk.fwd), _p(bias), _p(gamma), _p(beta), _p(z), z.stride(1), _p(stats),
              ya, B, L, C, float(drop_p), int(seed), _p(step), _stream())
    return y, (z, stats)


def deconv_bwd(dy, x, saved, pk, gamma, beta, dw, dbias, dgamma, dbeta, drop_p=0.0, seed=0, step=None):
    z, stats = saved
    ldx, B, L, C = _rows(x)
    lddy = _rows(dy)[0]
    dev = x.device
    dz = torch.empty(B, 2 * L, C, device=dev, dtype=torch.float32)
    dx = torch.empty(B, L, C, device=dev, dtype=torch.float32)
    _lib.call("oph_deconv_bwd", _p(dy), lddy, _act(x), _p(z), z.stride(1), _p(stats), _p(pk.bwd), _p(gamma),
              _p(beta), _p(dz), dz.stride(1), _p(dx), dx.stride(1), _p(dw), _p(dbias), _p(dgamma), _p(dbeta),
              B, L, C, float(drop_p), int(seed), _p(step), _stream())
    _park(dz)
    return dx


# ---------------------------------------------------------------------------------------------- embedding
def embed_fwd(ids, table):
    assert ids.dtype == torch.int32 and ids.is_contiguous()
    B, N = ids.shape
    E = table.shape[1]
    out = torch.empty(B, N, E, device=table.device, dtype=torch.float32)
    _lib.call("oph_embed_fwd", _p(ids), _p(table), _p(out), E, B * N, E, _stream())
    return out


def embed_bwd(ids, dout, dtable):
    ld = _rows(dout)[0]
    _lib.call("oph_embed_bwd", _p(ids), _p(dout), ld, _p(dtable), ids.numel(), dout.shape[2], _stream())


# ---------------------------------------------------------------------------------------------- attention
def split_planes(t, into=None):
    """Split-bf16 planes (hi, lo) [B, L, pad8(C)] of a [B, L, C] activation that did not come out of a row-wise kernel
    (mels, embeddings, fed K / V, the decoder-input gradient): one oph_split_planes launch.  `into` = (hi, lo) views to
    fill instead of new buffers."""
    ld, B, L, C = _rows(t)
    if into is None:
        ldp = _pad8(C)
        hi = torch.empty(B, L, ldp, device=t.device, dtype=torch.bfloat16)[:, :, :C]
        lo = torch.empty(B, L, ldp, device=t.device, dtype=torch.bfloat16)[:, :, :C]
    else:
        hi, lo = into
    assert hi.stride(1) == lo.stride(1) and hi.stride(2) == 1 and hi.stride(0) == L * hi.stride(1)
    _lib.call("oph_split_planes", _p(t), ld, B * L, C, _p(hi), _p(lo), hi.stride(1), _stream())
    return hi, lo


def ensure_planes(t, cache=False):
    """Planes of t: the ones attached by the producing kernel, else a fresh split.  cache=True attaches the result to the
    tensor: only for tensors whose contents do not change afterwards (K / V inside the autoregressive loop)."""
    pl = getattr(t, "_oph_planes", None)
    if pl is None:
        pl = split_planes(t)
        if cache:
            t._oph_planes = pl
    return pl


def _new_planes(t):
    """Uninitialised planes for an fp32 [B, T, ld >= N] scratch whose logical width is t.shape[2]."""
    B, L, C = t.shape
    ldp = _pad8(C)
    hi = torch.empty(B, L, ldp, device=t.device, dtype=torch.bfloat16)
    lo = torch.empty(B, L, ldp, device=t.device, dtype=torch.bfloat16)
    t._oph_planes = (hi[:, :, :C], lo[:, :, :C])


def _guide(gts, mse, extra=None):
    """`oph_guide` of the batch's own attention targets [B, Ng, Tg] (None -> analytic global guide).  Outside the block the
    guided loss sees 1.0 and the MSE variant 0.0 (architectures.py:263, 275).  extra = (col_g, col_h, c_aout): gradient
    inputs of the CDP / Ain / Aout terms from attention_extra_fwd (backward only)."""
    if gts is None and extra is None:
        return None
    gd = _lib.Guide()
    if gts is not None:
        assert gts.is_cuda and gts.dtype == torch.float32 and gts.dim() == 3 and gts.stride(2) == 1
        gd.w, gd.item_stride, gd.ld, gd.Ng, gd.Tg = gts.data_ptr(), gts.stride(0), gts.stride(1), gts.shape[1], gts.shape[2]
        gd.pad, gd.mse = (0.0 if mse else 1.0), int(bool(mse))
    if extra is not None:
        col_g, col_h, c_aout = extra
        gd.col_g, gd.col_h, gd.c_aout = col_g.data_ptr(), col_h.data_ptr(), float(c_aout)
    return ctypes.byref(gd)


def attention_guide_sum(A, att_acc, maxN, maxT, g, gts=None, mse=False):
    """att_acc += the guided-attention sum over externally supplied alignments A [B, T, N] (FixedAttention)."""
    B, T, N = A.shape
    assert A.is_cuda and A.dtype == torch.float32 and A.stride(2) == 1 and (B == 1 or A.stride(0) == T * A.stride(1))
    _lib.call("oph_attention_guide_sum", _p(A), A.stride(1), B, T, N, _p(att_acc), int(maxN), int(maxT), float(g),
              _guide(gts, mse), _stream())


def attention_extra_fwd(A, c_cdp, c_ain, acc3):
    """CDP / Ain / Aout sums of the alignments A [B, T, N] into acc3 (device double[3]); returns the per-key gradient
    factors (col_g, col_h) [B, N] for attention_bwd (architectures.py:283-321)."""
    B, T, N = A.shape
    assert A.stride(2) == 1 and A.stride(0) == T * A.stride(1)
    col_g = torch.empty(B, N, device=A.device, dtype=torch.float32)
    col_h = torch.empty(B, N, device=A.device, dtype=torch.float32)
    _lib.call("oph_attention_extra_fwd", _p(A), A.stride(1), B, T, N, float(c_cdp), float(c_ain), _p(col_g), _p(col_h),
              _p(acc3), _stream())
    return col_g, col_h


def attention_extra_finalize(acc3, comps8, B, T, N, w_cdp, w_ain, w_aout, add_to_total):
    _lib.call("oph_attention_extra_finalize", _p(acc3), _p(comps8), B, T, N, float(w_cdp), float(w_ain), float(w_aout),
              int(bool(add_to_total)), _stream())


def attention_fwd(Q, K, V, R=None, prev_max=None, win=3, want_alignments=False, want_argmax=True, att_acc=None,
                  maxN=1, maxT=1, g=0.2, gts=None, mse=False, need_A=True):
    """need_A=False (inference): the probabilities [B, T, N] are not wanted as such (the transposed `alignments` are a
    separate output).  The one-kernel attention (d == 256, N <= 256) then keeps them on chip; the three-launch path needs
    the buffer as scratch either way."""
    ldq, B, T, d = _rows(Q)
    ldk, _, N, _ = _rows(K)
    _rows(V)
    dev = Q.device
    ldA = _pad4(N)
    if R is None:
        R = torch.empty(B, T, d, device=dev, dtype=torch.float32)
    align = torch.empty(B, N, T, device=dev, dtype=torch.float32) if want_alignments else None
    argmax = torch.empty(B, T, device=dev, dtype=torch.int32) if want_argmax else None
    _rows(R)
    pq, pk_, pv = ensure_planes(Q), ensure_planes(K), ensure_planes(V)      # locals keep fresh splits alive over the call
    if need_A or not (d == 256 and N <= 256) or _lib.debug_flags() & (524288 | 8):
        A = torch.zeros(B, T, ldA, device=dev, dtype=torch.float32)[:, :, :N]
        _new_planes(A)
        Aa = _act(A)
    else:
        A, Aa = None, _lib.Act()
    _lib.call("oph_attention_fwd", _act(Q, planes=pq), _act(K, planes=pk_), _act(V, planes=pv), Aa, _act(R), _p(align),
              _p(argmax), _p(prev_max), int(win), _p(att_acc), int(maxN), int(maxT), float(g), B, T, N, d, _guide(gts, mse),
              _stream())
    return R, A, align, argmax


def attention_bwd(dR, Q, K, V, A, dq_addend=None, att_coef=0.0, maxN=1, maxT=1, g=0.2, dK=None, dV=None, gts=None,
                  mse=False, extra=None):
    ldq, B, T, d = _rows(Q)
    ldk, _, N, _ = _rows(K)
    dev = Q.device
    dA = torch.zeros(B, T, A.stride(1), device=dev, dtype=torch.float32)[:, :, :N]
    if dq_addend is not None:   # may be a strided view (second half of the [R,Q] gradient): mirror its layout
        _rows(dq_addend)
        dQ = torch.empty_strided(dq_addend.shape, dq_addend.stride(), device=dev, dtype=torch.float32)
    else:
        dQ = torch.empty(B, T, d, device=dev, dtype=torch.float32)
    if dK is None:
        dK = torch.empty(B, N, d, device=dev, dtype=torch.float32)
    if dV is None:
        dV = torch.empty(B, N, d, device=dev, dtype=torch.float32)
    _new_planes(dA)
    pr, pq, pk_, pv, pa = (ensure_planes(t) for t in (dR, Q, K, V, A))     # locals keep fresh splits alive over the call
    _lib.call("oph_attention_bwd", _act(dR, planes=pr), _act(Q, planes=pq), _act(K, planes=pk_), _act(V, planes=pv),
              _act(A, planes=pa), _act(dA),
              _p(dQ), dQ.stride(1), _p(dq_addend), dq_addend.stride(1) if dq_addend is not None else 0,
              _p(dK), dK.stride(1), _p(dV), dV.stride(1), float(att_coef), int(maxN), int(maxT), float(g),
              B, T, N, d, _guide(gts, mse, extra), _stream())
    return dQ, dK, dV


# ---------------------------------------------------------------------------------------------- losses / optimiser
def recon_loss(logits, target, acc, squash, w_l1, w_bd, w_l2, want_grad=True):
    ldl, B, L, C = _rows(logits)
    # the kernel pairs row r of the logits with row r of the target: a target of another length (features not padded
    # to r * T, a truncated .npy) would pair rows of different utterances -- TF raises a shape error here, so do we
    assert tuple(target.shape) == tuple(logits.shape), \
        "target %s does not match the predictions %s" % (tuple(target.shape), tuple(logits.shape))
    assert target.is_cuda and target.dtype == torch.float32 and target.stride(2) == 1
    assert B == 1 or target.stride(0) == L * target.stride(1), "target rows must be uniformly strided"
    ldt = target.stride(1)
    dl = new_act(B, L, C, logits.device) if want_grad else None
    _lib.call("oph_recon_loss", _p(logits), ldl, _p(target), ldt, _p(dl), dl.stride(1) if want_grad else 0,
              B * L, C, int(bool(squash)), float(w_l1), float(w_bd), float(w_l2), _p(acc), _stream())
    return dl


def loss_finalize(acc, out, n_recon, n_att, w_l1, w_bd, w_att, w_l2, has_att, squash):
    _lib.call("oph_loss_finalize", _p(acc), _p(out), float(n_recon), float(n_att), float(w_l1), float(w_bd),
              float(w_att), float(w_l2), int(bool(has_att)), int(bool(squash)), _stream())


def adam_prepare(global_step, lr_t, lr0, beta1, beta2, decay_lr, warmup=4000.0):
    _lib.call("oph_adam_prepare", _p(global_step), _p(lr_t), float(lr0), float(beta1), float(beta2),
              int(bool(decay_lr)), float(warmup), _stream())


def adam_clip(p, m, v, g, lr_t, beta1, beta2, eps, clip=1.0, grad_scale=1.0):
    _lib.call("oph_adam_clip", _p(p), _p(m), _p(v), _p(g), p.numel(), _p(lr_t), float(beta1), float(beta2),
              float(eps), float(clip), float(grad_scale), _stream())


def step_inc(global_step):
    _lib.call("oph_step_inc", _p(global_step), _stream())


# ---------------------------------------------------------------------------------------------- raw GEMM (tests)
def gemm_nt(A, Bm, b_mode=1, bias=None, alpha=1.0):
    """A [.., M, K] @ (Bm [.., N, K]^T if b_mode == 1 else Bm [.., K, N])."""
    batched = A.dim() == 3
    A3 = A if batched else A[None]
    B3 = Bm if batched else Bm[None]
    nb, M, K = A3.shape
    N = B3.shape[1] if b_mode == 1 else B3.shape[2]
    C = torch.zeros(nb, M, _pad4(N), device=A.device, dtype=torch.float32)
    _lib.call("oph_gemm_nt", _p(A3), A3.stride(1), _p(B3), B3.stride(1), _p(C), C.stride(1), _p(bias), M, N, K, b_mode,
              float(alpha), nb, A3.stride(0), B3.stride(0), C.stride(0), _stream())
    C = C[:, :, :N]
    return C if batched else C[0]


def gemm_tn(A, Bm, splits=1):
    """A [R, M]^T @ Bm [R, N] with split-K atomics."""
    R, M = A.shape
    N = Bm.shape[1]
    C = torch.zeros(M, _pad4(N), device=A.device, dtype=torch.float32)
    _lib.call("oph_gemm_tn", _p(A), A.stride(0), _p(Bm), Bm.stride(0), _p(C), C.stride(0), M, N, R, int(splits), _stream())
    return C[:, :N]
