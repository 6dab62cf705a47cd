"""Bring-up check of every C-ABI entry on a real B200 against the fp64 torch oracle (test tooling).

Prints one line per case (max-abs error vs fp64) and keeps going on failures so that a single gpurun call
gives the whole picture.  `python tests/gpu_check.py [groups] [--perf]`.
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ophelia_b200 import ops  # noqa: E402
from oracle import dctts_torch as ot  # noqa: E402

dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
torch.manual_seed(0)
RESULTS = []


def report(name, err, tol):
    ok = bool(err <= tol)
    RESULTS.append((name, err, tol, ok))
    print("%-58s err=%.3e tol=%.1e %s" % (name, err, tol, "ok" if ok else "FAIL"), flush=True)


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        RESULTS.append((name, float("nan"), 0, False))
        print("%-58s EXCEPTION %s" % (name, e), flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            print("CUDA context is broken: %s" % e2, flush=True)
            summary()
            sys.exit(2)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, dtype=torch.float64) * scale)


def f32(t):
    return t.to(torch.float32).to(dev).contiguous()


def act32(t):
    """[B, L, C] activation on the GPU with a row stride padded to 4 floats."""
    t = t.detach()
    out = ops.new_act(t.shape[0], t.shape[1], t.shape[2], dev)
    out.copy_(t.to(torch.float32))
    return out


def maxerr(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


# ------------------------------------------------------------------------------------------ raw GEMMs
def case_gemm_nt(M, N, K, b_mode, batch=0):
    def fn():
        sh = (batch,) if batch else ()
        A = rnd(*sh, M, K)
        Bm = rnd(*sh, N, K) if b_mode == 1 else rnd(*sh, K, N)
        bias = rnd(N)
        ref = A @ (Bm.transpose(-1, -2) if b_mode == 1 else Bm) * 0.5 + bias
        out = ops.gemm_nt(f32(A), f32(Bm), b_mode, f32(bias), 0.5)
        report("gemm_nt M%d N%d K%d mode%d batch%d" % (M, N, K, b_mode, batch), maxerr(out, ref), 2e-4 * np.sqrt(K))
    return fn


def case_gemm_tn(R, M, N, splits):
    def fn():
        A, Bm = rnd(R, M), rnd(R, N)
        ref = A.t() @ Bm
        out = ops.gemm_tn(f32(A), f32(Bm), splits)
        report("gemm_tn R%d M%d N%d splits%d" % (R, M, N, splits), maxerr(out, ref), 2e-4 * np.sqrt(R))
    return fn


# ------------------------------------------------------------------------------------------ layers
def conv_params(scope, k, cin, cout, hc=False, deconv=False):
    P = {}
    if deconv:
        P[scope + "/conv2d_transpose/kernel"] = rnd(1, 3, cout, cin, scale=(2.6 / (3 * cout)) ** 0.5)
        P[scope + "/conv2d_transpose/bias"] = rnd(cout, scale=0.2)
    else:
        P[scope + "/conv1d/kernel"] = rnd(k, cin, cout, scale=(2.6 / (k * cin)) ** 0.5)
        P[scope + "/conv1d/bias"] = rnd(cout, scale=0.2)
    names = ["H1", "H2"] if hc else ["normalize"]
    c = cout // 2 if hc else cout
    for n in names:
        P[scope + "/%s/gamma" % n] = 1.0 + rnd(c, scale=0.2)
        P[scope + "/%s/beta" % n] = rnd(c, scale=0.2)
    for v in P.values():
        v.requires_grad_(True)
    return P


def case_conv1d(B, L, cin, cout, act, padding, in_shift=0, norm=True):
    def fn():
        P = conv_params("c", 1, cin, cout)
        x = rnd(B, L, cin).requires_grad_(True)
        xs = x
        if in_shift:
            xs = torch.cat([torch.zeros_like(x[:, :in_shift]), x[:, :-in_shift]], 1)
        ref = ot.conv1d(P, xs, "c", 1, 1, "CAUSAL" if padding else "SAME", "relu" if act else None,
                        "layer" if norm else None)
        dy = rnd(B, L, cout)
        if act:   # a ReLU unit within fp32 noise of its kink may take the other branch: give those no gradient
            with torch.no_grad():
                pre = ot.conv1d(P, xs, "c", 1, 1, "CAUSAL" if padding else "SAME", None, "layer" if norm else None)
                dy[pre.abs() < 1e-3] = 0.0
        ref.backward(dy)
        w = f32(P["c/conv1d/kernel"])
        pk = ops.PackedConv(w)
        bias, gamma, beta = f32(P["c/conv1d/bias"]), f32(P["c/normalize/gamma"]), f32(P["c/normalize/beta"])
        xg = act32(x)
        y, ysig, saved = ops.conv1d_fwd(xg, pk, bias, gamma, beta, 1, padding, in_shift, act, norm, save=True,
                                        want_sigmoid=True)
        tag = "conv1d %dx%d %d->%d act%d pad%d sh%d n%d" % (B, L, cin, cout, act, padding, in_shift, norm)
        report(tag + " fwd", maxerr(y, ref), 2e-4)
        dw, db = torch.zeros_like(w), torch.zeros_like(bias)
        dg, dbe = torch.zeros_like(gamma), torch.zeros_like(beta)
        dx = ops.conv1d_bwd(act32(dy), xg, saved, pk, gamma, beta, dw, db, dg, dbe, 1, padding, in_shift, act, norm)
        report(tag + " dx", maxerr(dx, x.grad), 5e-4)
        e = (dx.detach().double().cpu() - x.grad).abs().amax(-1)
        if float(e.max()) > 5e-4:
            print("   bad dx rows (b,t):", [(int(i[0]), int(i[1]), float(e[i[0], i[1]])) for i in (e > 5e-4).nonzero()[:12]])
        gw = P["c/conv1d/kernel"].grad
        report(tag + " dw", maxerr(dw, gw), 1e-3 * max(1.0, float(gw.abs().max())))
        report(tag + " dbias", maxerr(db, P["c/conv1d/bias"].grad), 2e-3)
        if norm:
            report(tag + " dgamma", maxerr(dg, P["c/normalize/gamma"].grad), 2e-3)
            report(tag + " dbeta", maxerr(dbe, P["c/normalize/beta"].grad), 2e-3)
    return fn


def case_hc(B, L, C, k, rate, padding):
    def fn():
        P = conv_params("h", k, C, 2 * C, hc=True)
        x = rnd(B, L, C).requires_grad_(True)
        ref = ot.hc(P, x, "h", k, rate, "CAUSAL" if padding else "SAME")
        dy = rnd(B, L, C)
        ref.backward(dy)
        w = f32(P["h/conv1d/kernel"])
        pk = ops.PackedConv(w)
        prm = [f32(P[n]) for n in ("h/conv1d/bias", "h/H1/gamma", "h/H1/beta", "h/H2/gamma", "h/H2/beta")]
        bias, g1, b1, g2, b2 = prm
        xg = f32(x)
        y, saved = ops.hc_fwd(xg, pk, bias, g1, b1, g2, b2, rate, padding, True, save=True)
        tag = "hc %dx%d C%d k%d r%d pad%d" % (B, L, C, k, rate, padding)
        report(tag + " fwd", maxerr(y, ref), 2e-4)
        grads = [torch.zeros_like(t) for t in [w] + prm]
        dx = ops.hc_bwd(f32(dy), xg, saved, pk, g1, b1, g2, b2, grads[0], grads[1], grads[2], grads[3], grads[4],
                        grads[5], rate, padding, True)
        report(tag + " dx", maxerr(dx, x.grad), 5e-4)
        gw = P["h/conv1d/kernel"].grad
        report(tag + " dw", maxerr(grads[0], gw), 1e-3 * max(1.0, float(gw.abs().max())))
        for i, n in enumerate(("h/conv1d/bias", "h/H1/gamma", "h/H1/beta", "h/H2/gamma", "h/H2/beta")):
            report(tag + " d" + n.split("/", 1)[1], maxerr(grads[i + 1], P[n].grad), 2e-3)
    return fn


def case_hc_planes(B, L, C, k, rate, padding):
    """Second layer of a chain: its input arrives with split-bf16 planes (copied, not converted, by the GEMMs)."""
    def fn():
        P0 = conv_params("p", 1, C, C)
        P = conv_params("h", k, C, 2 * C, hc=True)
        x0 = act32(rnd(B, L, C))
        pk0 = ops.PackedConv(f32(P0["p/conv1d/kernel"]))
        x, _, _ = ops.conv1d_fwd(x0, pk0, f32(P0["p/conv1d/bias"]), f32(P0["p/normalize/gamma"]), f32(P0["p/normalize/beta"]))
        assert getattr(x, "_oph_planes", None) is not None
        hi, lo = x._oph_planes
        report("planes hi+lo == fp32 (C%d)" % C, maxerr(hi.float() + lo.float(), x), 2e-5 * float(x.abs().max()))
        xr = x.detach().double().cpu().requires_grad_(True)
        ref = ot.hc(P, xr, "h", k, rate, "CAUSAL" if padding else "SAME")
        dy = rnd(B, L, C)
        ref.backward(dy)
        w = f32(P["h/conv1d/kernel"])
        pk = ops.PackedConv(w)
        prm = [f32(P[n]) for n in ("h/conv1d/bias", "h/H1/gamma", "h/H1/beta", "h/H2/gamma", "h/H2/beta")]
        bias, g1, b1, g2, b2 = prm
        y, saved = ops.hc_fwd(x, pk, bias, g1, b1, g2, b2, rate, padding, True, save=True)
        tag = "hc(planes in) %dx%d C%d k%d r%d pad%d" % (B, L, C, k, rate, padding)
        report(tag + " fwd", maxerr(y, ref), 2e-4)
        grads = [torch.zeros_like(t) for t in [w] + prm]
        dx = ops.hc_bwd(f32(dy), x, saved, pk, g1, b1, g2, b2, grads[0], grads[1], grads[2], grads[3], grads[4],
                        grads[5], rate, padding, True)
        report(tag + " dx", maxerr(dx, xr.grad), 5e-4)
        gw = P["h/conv1d/kernel"].grad
        report(tag + " dw", maxerr(grads[0], gw), 1e-3 * max(1.0, float(gw.abs().max())))
        report(tag + " dbias", maxerr(grads[1], P["h/conv1d/bias"].grad), 2e-3)
    return fn


def case_deconv(B, L, C):
    def fn():
        P = conv_params("d", 3, C, C, deconv=True)
        x = rnd(B, L, C).requires_grad_(True)
        ref = ot.conv1d_transpose(P, x, "d")
        dy = rnd(B, 2 * L, C)
        ref.backward(dy)
        w = f32(P["d/conv2d_transpose/kernel"])
        pk = ops.PackedConv(w, deconv=True)
        bias, gamma, beta = f32(P["d/conv2d_transpose/bias"]), f32(P["d/normalize/gamma"]), f32(P["d/normalize/beta"])
        xg = f32(x)
        y, saved = ops.deconv_fwd(xg, pk, bias, gamma, beta, save=True)
        tag = "deconv %dx%d C%d" % (B, L, C)
        report(tag + " fwd", maxerr(y, ref), 2e-4)
        dw, db, dg, dbe = (torch.zeros_like(t) for t in (w, bias, gamma, beta))
        dx = ops.deconv_bwd(f32(dy), xg, saved, pk, gamma, beta, dw, db, dg, dbe)
        report(tag + " dx", maxerr(dx, x.grad), 5e-4)
        gw = P["d/conv2d_transpose/kernel"].grad
        report(tag + " dw", maxerr(dw, gw), 1e-3 * max(1.0, float(gw.abs().max())))
        report(tag + " dbias", maxerr(db, P["d/conv2d_transpose/bias"].grad), 2e-3)
        report(tag + " dgamma", maxerr(dg, P["d/normalize/gamma"].grad), 2e-3)
    return fn


class _HP(object):
    d = 256
    attention_win_size = 3
    concatenate_query = True
    g = 0.2


def case_attention(B, T, N, mono):
    def fn():
        hp = _HP()
        hp.max_N, hp.max_T = N, T
        d = 256
        Q, K, V = rnd(B, T, d).requires_grad_(True), rnd(B, N, d).requires_grad_(True), rnd(B, N, d).requires_grad_(True)
        prev = torch.randint(0, N - 2, (B,)) if mono else None
        Rref, Aref, mxref = ot.Attention(hp, Q, K, V, mono, prev)
        W = ot.attention_guide(hp, torch.float64)
        att = (Aref * W[None]).sum() / float(B * N * T)
        dRp = rnd(B, T, 2 * d)
        lw = 0.3333
        ((Rref * dRp).sum() + lw * att).backward()
        buf = torch.zeros(B, T, 2 * d, device=dev)      # [R | Q] buffer
        Qg = buf[:, :, d:]
        Qg.copy_(f32(Q))
        KV = torch.zeros(B, N, 2 * d, device=dev)
        KV[:, :, :d] = f32(K)
        KV[:, :, d:] = f32(V)
        Kg, Vg = KV[:, :, :d], KV[:, :, d:]
        acc = torch.zeros(4, device=dev, dtype=torch.float64)
        prevg = prev.to(torch.int32).to(dev) if mono else None
        R, A, align, argmax = ops.attention_fwd(Qg, Kg, Vg, R=buf[:, :, :d], prev_max=prevg, win=3,
                                                want_alignments=True, att_acc=acc[3:], maxN=N, maxT=T, g=0.2)
        tag = "attention B%d T%d N%d mono%d" % (B, T, N, mono)
        report(tag + " R'", maxerr(buf, Rref), 2e-4)
        report(tag + " alignments", maxerr(align, Aref), 1e-4)
        report(tag + " argmax(mismatches)", float((argmax.cpu().long() != mxref).sum()), 0)
        report(tag + " att_loss", abs(float(acc[3]) / (B * N * T) - float(att.detach())), 1e-6)
        dRg = f32(dRp)
        dKV = torch.zeros(B, N, 2 * d, device=dev)
        dQ, dK, dV = ops.attention_bwd(dRg[:, :, :d], Qg, Kg, Vg, A, dq_addend=dRg[:, :, d:],
                                       att_coef=lw / (B * N * T), maxN=N, maxT=T, g=0.2,
                                       dK=dKV[:, :, :d], dV=dKV[:, :, d:])
        report(tag + " dQ", maxerr(dQ, Q.grad), 5e-4)
        report(tag + " dK", maxerr(dK, K.grad), 5e-4 * max(1.0, float(K.grad.abs().max())))
        report(tag + " dV", maxerr(dV, V.grad), 5e-4 * max(1.0, float(V.grad.abs().max())))
    return fn


def case_embed():
    ids = torch.randint(0, 65, (3, 50), dtype=torch.int32)
    ids[:, 40:] = 0
    table = rnd(65, 128)
    ref = table.clone()
    ref[0] = 0
    out = ops.embed_fwd(ids.to(dev), f32(table))
    report("embed fwd", maxerr(out, ref[ids.long()]), 1e-6)
    dout = rnd(3, 50, 128)
    dt = torch.zeros(65, 128, device=dev)
    ops.embed_bwd(ids.to(dev), f32(dout), dt)
    refg = torch.zeros(65, 128, dtype=torch.float64)
    refg.index_add_(0, ids.long().reshape(-1), dout.reshape(-1, 128))
    refg[0] = 0
    report("embed bwd", maxerr(dt, refg), 1e-5)


def case_loss_adam():
    B, L, C = 3, 70, 80
    logits = rnd(B, L, C).requires_grad_(True)
    tgt = torch.rand(B, L, C, dtype=torch.float64)
    Y = torch.sigmoid(logits)
    l1 = (Y - tgt).abs().mean()
    bd = torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt)
    l2 = ((Y - tgt) ** 2).mean()
    (0.4 * l1 + 0.5 * bd + 0.1 * l2).backward()
    acc = torch.zeros(4, device=dev, dtype=torch.float64)
    dl = ops.recon_loss(f32(logits), f32(tgt), acc, True, 0.4, 0.5, 0.1)
    out = torch.zeros(5, device=dev)
    ops.loss_finalize(acc, out, B * L * C, 1.0, 0.4, 0.5, 0.0, 0.1, True, True)
    report("recon_loss dlogits", maxerr(dl, logits.grad), 1e-8)
    ref = torch.tensor([float(0.4 * l1 + 0.5 * bd + 0.1 * l2), float(l1), float(bd), 0.0, float(l2)])
    report("loss components", maxerr(out, ref), 2e-6)
    # adam
    n = 1003
    p, g = rnd(n), rnd(n, scale=2.0)
    m, v = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    pg, mg, vg, gg = f32(p), f32(m), f32(v), f32(g)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    lr_t = torch.zeros(2, device=dev)
    from oracle import dctts_numpy as on
    pn, mn, vn = p.numpy(), m.numpy(), v.numpy()
    for t in range(1, 4):
        lr = on.learning_rate_decay(0.001, t - 1)
        pn, mn, vn = on.adam_step(pn, mn, vn, g.numpy() * 0.5, t, lr)
        ops.adam_prepare(step, lr_t, 0.001, 0.9, 0.999, True)
        ops.adam_clip(pg, mg, vg, gg, lr_t, 0.9, 0.999, 1e-8, 1.0, 0.5)
        ops.step_inc(step)
    report("adam 3 steps", maxerr(pg, torch.tensor(pn)), 1e-6)
    report("global_step", abs(int(step.item()) - 3), 0)


def case_pack_batch():
    """One-launch re-pack (oph_pack_plan_add / oph_pack_run) must reproduce oph_conv_pack bit for bit."""
    import ctypes
    from ophelia_b200 import _lib
    lib = _lib.load()
    ws = [(torch.randn(3, 256, 512, device=dev), False), (torch.randn(1, 80, 256, device=dev), False),
          (torch.randn(1, 513, 513, device=dev), False), (torch.randn(1, 3, 512, 512, device=dev), True)]
    pks = [ops.PackedConv(w, deconv=d) for w, d in ws]
    ref = [(pk.fwd.clone(), pk.bwd.clone()) for pk in pks]
    for pk in pks:
        pk.fwd.zero_(); pk.bwd.zero_()
    jb = int(lib.oph_pack_job_bytes())
    host = ctypes.create_string_buffer(3 * len(pks) * jb)
    nj, nb = ctypes.c_int(0), ctypes.c_longlong(0)
    for pk in pks:
        _lib.call("oph_pack_plan_add", ctypes.cast(host, ctypes.c_void_p), 3 * len(pks), ctypes.byref(nj), ctypes.byref(nb),
                  pk.w.data_ptr(), pk.k, pk.cin, pk.cout, int(pk.deconv), pk.fwd.data_ptr(), pk.bwd.data_ptr())
    plan = torch.frombuffer(bytearray(host.raw[:nj.value * jb]), dtype=torch.uint8).to(dev)
    _lib.call("oph_pack_run", plan.data_ptr(), nj.value, nb.value, torch.cuda.current_stream().cuda_stream)
    bad = sum(int((pk.fwd != r[0]).sum()) + int((pk.bwd != r[1]).sum()) for pk, r in zip(pks, ref))
    report("pack_run == conv_pack (mismatching bytes, %d jobs)" % nj.value, float(bad), 0)


def case_dropout():
    """Dropout (modules.py:141,205): element-wise keep-prob 1-rate, kept values scaled by 1/(1-rate), the backward
    pass re-creates the same mask from (seed, global_step, element index)."""
    B, L, C, rate = 4, 300, 256, 0.05
    P = conv_params("h", 3, C, 2 * C, hc=True)
    w = f32(P["h/conv1d/kernel"]); pk = ops.PackedConv(w)
    prm = [f32(P[n]) for n in ("h/conv1d/bias", "h/H1/gamma", "h/H1/beta", "h/H2/gamma", "h/H2/beta")]
    x = f32(rnd(B, L, C))
    step = torch.full((1,), 7, dtype=torch.int64, device=dev)
    y0, _ = ops.hc_fwd(x, pk, *prm, rate=3, padding=1, norm=True)
    y1, saved = ops.hc_fwd(x, pk, *prm, rate=3, padding=1, norm=True, drop_p=rate, seed=1234, step=step, save=True)
    y2, _ = ops.hc_fwd(x, pk, *prm, rate=3, padding=1, norm=True, drop_p=rate, seed=1234, step=step)
    step2 = torch.full((1,), 8, dtype=torch.int64, device=dev)
    y3, _ = ops.hc_fwd(x, pk, *prm, rate=3, padding=1, norm=True, drop_p=rate, seed=1234, step=step2)
    keep = (y1 != 0)
    frac = float(keep.float().mean())
    report("dropout keep fraction (rate %.2f)" % rate, abs(frac - (1 - rate)), 4 * (rate * (1 - rate) / keep.numel()) ** 0.5 + 1e-4)
    report("dropout scale 1/(1-rate)", maxerr(y1[keep], y0[keep] / (1 - rate)), 1e-5)
    report("dropout mask repeatable for one step", float((y1 != y2).sum()), 0)
    report("dropout mask changes with global_step (same fraction kept)", abs(float(((y3 != 0) == keep).float().mean()) - (1 - 2 * rate * (1 - rate))), 5e-3)
    # backward uses the same mask: dx(dropout) == dx(no dropout, dy * mask / (1-rate))
    dy = f32(rnd(B, L, C))
    g = [torch.zeros_like(t) for t in [w] + prm]
    dx1 = ops.hc_bwd(dy, x, saved, pk, *prm[1:], g[0], g[1], g[2], g[3], g[4], g[5], 3, 1, True, drop_p=rate, seed=1234, step=step)
    g2 = [torch.zeros_like(t) for t in [w] + prm]
    dym = (dy * keep.float() / (1 - rate)).contiguous()
    dx0 = ops.hc_bwd(dym, x, saved, pk, *prm[1:], g2[0], g2[1], g2[2], g2[3], g2[4], g2[5], 3, 1, True)
    report("dropout backward mask == forward mask (dx)", maxerr(dx1, dx0), 1e-5)
    report("dropout backward mask == forward mask (dw)", maxerr(g[0], g2[0]), 1e-3)
    # conv1d (tail fused into the GEMM epilogue for Cout <= 256): forward mask must be the one ln_act_bwd re-creates
    Pc = conv_params("c", 1, C, C)
    wc = f32(Pc["c/conv1d/kernel"]); pkc = ops.PackedConv(wc)
    bc, gc, bec = f32(Pc["c/conv1d/bias"]), f32(Pc["c/normalize/gamma"]), f32(Pc["c/normalize/beta"])
    c0, _, _ = ops.conv1d_fwd(x, pkc, bc, gc, bec, 1, 1, 0, 0, True)
    c1, _, savedc = ops.conv1d_fwd(x, pkc, bc, gc, bec, 1, 1, 0, 0, True, drop_p=rate, seed=77, step=step, save=True)
    keepc = (c1 != 0)
    report("conv1d dropout keep fraction", abs(float(keepc.float().mean()) - (1 - rate)), 4 * (rate * (1 - rate) / keepc.numel()) ** 0.5 + 1e-4)
    report("conv1d dropout scale", maxerr(c1[keepc], c0[keepc] / (1 - rate)), 1e-5)
    gr = [torch.zeros_like(t) for t in (wc, bc, gc, bec)]
    dxa = ops.conv1d_bwd(dy, x, savedc, pkc, gc, bec, gr[0], gr[1], gr[2], gr[3], 1, 1, 0, 0, True, drop_p=rate, seed=77, step=step)
    gr2 = [torch.zeros_like(t) for t in (wc, bc, gc, bec)]
    dymc = (dy * keepc.float() / (1 - rate)).contiguous()
    dxb = ops.conv1d_bwd(dymc, x, savedc, pkc, gc, bec, gr2[0], gr2[1], gr2[2], gr2[3], 1, 1, 0, 0, True)
    report("conv1d dropout backward mask == forward mask (dx)", maxerr(dxa, dxb), 1e-5)


def perf():
    """First timing of the dominant layer shape (AudioEnc/AudioDec highway conv at B=32, T=870, C=256)."""
    for (B, L, C, k) in [(32, 870, 256, 3), (32, 180, 512, 3), (64, 870, 256, 3)]:
        P = conv_params("h", k, C, 2 * C, hc=True)
        w = f32(P["h/conv1d/kernel"])
        pk = ops.PackedConv(w)
        prm = [f32(P[n]) for n in ("h/conv1d/bias", "h/H1/gamma", "h/H1/beta", "h/H2/gamma", "h/H2/beta")]
        x = torch.randn(B, L, C, device=dev)
        y = torch.empty_like(x)
        for _ in range(3):
            ops.hc_fwd(x, pk, *prm, rate=3, padding=1, y=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.hc_fwd(x, pk, *prm, rate=3, padding=1, y=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2.0 * k * C * 2 * C * B * L
        print("perf hc_fwd B%d L%d C%d k%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (B, L, C, k, ms, fl / ms / 1e9), flush=True)


def summary():
    bad = [r for r in RESULTS if not r[3]]
    print("==== %d cases, %d failed" % (len(RESULTS), len(bad)))
    for r in bad:
        print("FAILED:", r[0], r[1], r[2])


GROUPS = {
    "gemm_nt": [case_gemm_nt(*a) for a in [(128, 256, 64, 1, 0), (200, 180, 256, 1, 0), (870, 180, 256, 1, 3),
                                          (1000, 512, 192, 1, 0), (130, 80, 80, 1, 0), (300, 1025, 128, 1, 0),
                                          (200, 256, 180, 2, 0), (870, 256, 180, 2, 2), (64, 512, 60, 2, 0)]],
    "gemm_tn": [case_gemm_tn(*a) for a in [(1000, 256, 512, 3), (870, 180, 256, 1), (5000, 80, 256, 4),
                                          (400, 516, 1028, 2)]],
    "conv1d": [case_conv1d(2, 200, 80, 256, 1, 1, in_shift=1), case_conv1d(2, 200, 256, 80, 0, 1),
               case_conv1d(2, 60, 128, 512, 1, 0), case_conv1d(1, 150, 1024, 513, 0, 0),
               case_conv1d(1, 150, 513, 513, 1, 0), case_conv1d(2, 100, 256, 256, 0, 1, norm=False),
               case_conv1d(1, 37, 1025, 1025, 1, 0), case_conv1d(1, 150, 512, 1024, 1, 0)],
    "hc": [case_hc(*a) for a in [(2, 200, 256, 3, 1, 1), (2, 200, 256, 3, 27, 1), (2, 60, 512, 3, 9, 0),
                                 (2, 60, 512, 1, 1, 0), (1, 130, 1024, 3, 1, 0), (3, 129, 256, 3, 3, 0)]],
    "hc_planes": [case_hc_planes(2, 200, 256, 3, 3, 1), case_hc_planes(2, 70, 512, 3, 27, 0), case_hc_planes(1, 130, 1024, 3, 1, 0)],
    "deconv": [case_deconv(2, 50, 512), case_deconv(1, 131, 256)],
    "attention": [case_attention(2, 200, 60, False), case_attention(2, 210, 180, True),
                  case_attention(3, 130, 47, False), case_attention(2, 300, 256, True), case_attention(1, 129, 130, False),
                  case_attention(4, 85, 150, True), case_attention(3, 870, 180, False)],
    "misc": [case_embed, case_loss_adam, case_pack_batch, case_dropout],
}


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    only = [a for a in sys.argv[1:] if not a.startswith("-")]       # optional group names
    for name, cases in GROUPS.items():
        if only and name not in only:
            continue
        for fn in cases:
            run(name, fn)
    print("checks took %.1fs" % (time.time() - t0), flush=True)
    if "--perf" in sys.argv:
        run("perf", perf)
    summary()
    sys.exit(1 if any(not r[3] for r in RESULTS) else 0)


if __name__ == "__main__":
    main()
