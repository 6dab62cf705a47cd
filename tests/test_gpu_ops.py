"""Per-operator parity of every C-ABI entry (forward and backward) against the fp64 torch oracle.

The cases live in tests/gpu_check.py (also runnable stand-alone for bring-up); shapes cover non-multiples of the
128-row tile, N/K tails (80, 180, 513, 1025 channels), all dilation rates, causal/SAME, strided [R|Q] / [K|V]
buffers, the monotonic window mask, split-K weight gradients and batched attention GEMMs."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["gemm_nt", "gemm_tn", "conv1d", "hc", "hc_planes", "deconv", "attention", "misc"])
def test_ops_match_oracle(group):
    import gpu_check
    del gpu_check.RESULTS[:]
    for fn in gpu_check.GROUPS[group]:
        fn()
    bad = [r for r in gpu_check.RESULTS if not r[3]]
    assert gpu_check.RESULTS and not bad, bad


def test_fused_conv_tail_matches_oracle():
    """Optional epilogue mode of the GEMM (oph_gemm_debug_flags bit 65536): LayerNorm / ReLU / dropout / operand planes of
    conv layers with <= 256 output channels computed in the epilogue instead of a separate launch."""
    import gpu_check
    from ophelia_b200 import _lib
    lib = _lib.load()
    lib.oph_gemm_debug_flags(65536)
    try:
        n0 = lib.oph_launch_count()
        del gpu_check.RESULTS[:]
        for fn in gpu_check.GROUPS["conv1d"] + [gpu_check.case_dropout]:
            fn()
        bad = [r for r in gpu_check.RESULTS if not r[3]]
        assert gpu_check.RESULTS and not bad, bad
    finally:
        lib.oph_gemm_debug_flags(0)
    n_fused = lib.oph_launch_count() - n0
    del gpu_check.RESULTS[:]
    n1 = lib.oph_launch_count()
    for fn in gpu_check.GROUPS["conv1d"] + [gpu_check.case_dropout]:
        fn()
    assert lib.oph_launch_count() - n1 > n_fused            # the fused mode really skipped the separate tail launches


@pytest.mark.parametrize("C", [80, 256, 513])
def test_standalone_normalize_matches_oracle(C):
    """modules.normalize called directly (modules.py:47-75): layer norm over the channels with eps 1e-12 and biased
    variance, forward and backward (dx, dgamma, dbeta) against fp64; normtype None is the identity."""
    import numpy as np
    import torch
    from ophelia_b200 import modules, ops
    from ophelia_b200.variables import VariableStore, use_store
    from oracle import dctts_torch as ot
    torch.manual_seed(C)
    B, L = 3, 37
    x64 = (torch.randn(B, L, C, dtype=torch.float64) * 1.7 + 0.3).requires_grad_(True)
    gamma = (1.0 + 0.3 * torch.randn(C, dtype=torch.float64)).requires_grad_(True)
    beta = (0.2 * torch.randn(C, dtype=torch.float64)).requires_grad_(True)
    dy64 = torch.randn(B, L, C, dtype=torch.float64)
    y64 = ot.layer_norm({"n/gamma": gamma, "n/beta": beta}, x64, "n")
    y64.backward(dy64)
    store = VariableStore("cuda:0")
    x = ops.new_act(B, L, C, "cuda:0")
    x.copy_(x64.detach().float())
    with use_store(store):
        assert modules.normalize(x, scope="n", normtype=None) is x
        y = modules.normalize(x, scope="n")            # declares n/gamma, n/beta
        store.load_state_dict({"n/gamma": gamma.detach().numpy(), "n/beta": beta.detach().numpy()})
        store.grad_flat.zero_()
        with modules.Tape() as tape:
            y = modules.normalize(x, scope="n")
        dy = ops.new_act(B, L, C, "cuda:0")
        dy.copy_(dy64.float())
        dx = tape.backward(dy)
    err = lambda a, b: float((a.double().cpu() - b.detach()).abs().max())
    assert err(y, y64) < 2e-5
    assert err(dx, x64.grad) < 5e-5 * float(x64.grad.abs().max())  + 1e-6
    assert err(store.grad("n/gamma"), gamma.grad) < 1e-4 * float(gamma.grad.abs().max())
    assert err(store.grad("n/beta"), beta.grad) < 1e-4 * float(beta.grad.abs().max())


@pytest.mark.parametrize("C,rows,k,rate,causal,drop", [
    (256, 27840, 3, 3, True, 0.05),      # BASELINE shape: 109 row-tile pairs = 74 fused + 35 plain (two launches)
    (256, 74 * 256, 3, 27, False, 0.0),  # exactly one round of super-units, everything fused (one launch)
    (256, 60 * 256 + 37, 1, 1, True, 0.05),   # partial round that still fuses, ragged last tile
    (512, 19000, 3, 9, False, 0.05),     # 4 column blocks per super-unit, two warps per row in the tail
    (1024, 14000, 3, 1, False, 0.0),     # 8 column blocks, four warps per row
])
def test_fused_highway_tail_equals_separate_tail(C, rows, k, rate, causal, drop):
    """modules.hc with the highway tail inside the conv launch (the super-unit schedule of gemm_tc.cuh) against the same
    layer with the tail as its own launch (oph_gemm_debug_flags bit 131072; that path is pinned to the oracle by the `hc`
    group above and by the network tests): same arithmetic in the same order, so y, its operand planes, z and the row
    statistics must be bit-identical -- at sizes where whole rounds of CTA pairs run super-units."""
    import torch
    from ophelia_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(C + rows)
    dev = "cuda:0"
    B, L = (32, rows // 32) if rows % 32 == 0 else (1, rows)
    x = torch.randn(B, L, C, device=dev)
    x._oph_planes = ops.split_planes(x)
    w = torch.randn(k, C, 2 * C, device=dev) * (2.6 / (k * C)) ** 0.5
    pk = ops.PackedConv(w)
    bias = 0.1 * torch.randn(2 * C, device=dev)
    g1, g2 = (1.0 + 0.2 * torch.randn(C, device=dev) for _ in range(2))
    b1, b2 = (0.2 * torch.randn(C, device=dev) for _ in range(2))
    step = torch.tensor([3], dtype=torch.int64, device=dev)

    def run():
        n0 = lib.oph_launch_count()
        y, (z, stats) = ops.hc_fwd(x, pk, bias, g1, b1, g2, b2, rate, ops.CAUSAL if causal else ops.SAME, True, drop, 1234,
                                   step, save=True)
        torch.cuda.synchronize()
        return y, y._oph_planes, z, stats, lib.oph_launch_count() - n0
    # (flag 4: every CTA pair walks the k-blocks in the same order -- the two schedules give a row tile to different
    # pairs, and the default per-pair rotation of the k-block order would change the summation order of z)
    lib.oph_gemm_debug_flags(4 | 262144)          # (262144: the mixed schedule for a sparsely filled last round as well)
    try:
        y1, p1, z1, s1, n_fused = run()
        lib.oph_gemm_debug_flags(131072 | 4)
        y0, p0, z0, s0, n_sep = run()
    finally:
        lib.oph_gemm_debug_flags(0)
    assert n_sep == 2
    full_rounds_only = (-(-rows // 256)) % 74 == 0 or ((-(-rows // 256)) % 74) * 10 >= 74 * 7
    assert n_fused == (1 if full_rounds_only else 2), n_fused
    assert torch.equal(z1, z0) and torch.equal(s1, s0)
    assert torch.equal(y1, y0), float((y1 - y0).abs().max())
    assert torch.equal(p1[0], p0[0]) and torch.equal(p1[1], p0[1])
    assert torch.isfinite(y1).all() and float(y1.abs().max()) > 0


@pytest.mark.parametrize("B,T,N,mono", [(32, 870, 180, False), (10, 85, 150, True), (3, 200, 256, True), (2, 131, 17, False)])
def test_fused_attention_equals_three_launch_path(B, T, N, mono):
    """networks.Attention forward as one kernel (S and P stay in tensor / shared memory) against the three-launch path
    (two GEMM launches + the softmax kernel, oph_gemm_debug_flags bit 524288), which the `attention` group pins to the
    oracle: the scores come out of the same MMA sequence, so the argmax must agree exactly, the probabilities, the context
    vectors and the loss sum up to the summation order of the softmax; and it really is one launch."""
    import torch
    from ophelia_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(B * T + N)
    dev, d = "cuda:0", 256
    buf = torch.zeros(B, T, 2 * d, device=dev)
    Q = buf[:, :, d:]
    Q.copy_(torch.randn(B, T, d, device=dev))
    KV = torch.randn(B, N, 2 * d, device=dev)
    K, V = KV[:, :, :d], KV[:, :, d:]
    prev = torch.randint(0, max(1, N - 2), (B,), dtype=torch.int32, device=dev) if mono else None

    def run(need_A):
        acc = torch.zeros(1, device=dev, dtype=torch.float64)
        R = torch.zeros(B, T, d, device=dev)
        n0 = lib.oph_launch_count()
        R, A, align, argmax = ops.attention_fwd(Q, K, V, R=R, prev_max=prev, win=3, want_alignments=True, att_acc=acc,
                                                maxN=N, maxT=T, g=0.2, need_A=need_A)
        torch.cuda.synchronize()
        return R, A, align, argmax, float(acc[0]), lib.oph_launch_count() - n0
    R1, A1, al1, am1, acc1, n1 = run(True)
    R2, A2, al2, am2, acc2, n2 = run(False)          # inference form: the probabilities stay on chip
    _lib.set_debug_flags(524288)
    try:
        R0, A0, al0, am0, acc0, n0 = run(True)
    finally:
        _lib.set_debug_flags(0)
    # launches: Q / K / V planes are split by ensure_planes (3 small launches) in both modes
    assert n0 - n1 == 2 and n2 == n1 and A2 is None
    assert torch.equal(am1, am0) and torch.equal(am2, am0)
    assert float((al1 - al0).abs().max()) < 2e-6 and float((A1 - A0).abs().max()) < 2e-6
    assert torch.equal(al1.transpose(1, 2), A1)
    assert float((R1 - R0).abs().max()) < 2e-5 * max(1.0, float(R0.abs().max()))
    assert torch.equal(R2, R1) and torch.equal(al2, al1)
    assert abs(acc1 - acc0) < 1e-5 * max(1.0, abs(acc0))
    pl1, pl0 = A1._oph_planes, A0._oph_planes
    assert float((pl1[0].float() + pl1[1].float() - A1).abs().max()) < 1e-5
    assert float((pl0[0].float() + pl0[1].float() - A0).abs().max()) < 1e-5


def test_trapped_barrier_wait_leaves_a_decodable_record():
    """Every in-kernel barrier wait is bounded (a protocol bug traps instead of hanging the GPU).  The waiting thread
    leaves a record in mapped host memory that outlives the failed context; the library decodes it into the error message.
    The self-test kernel records a fake wait of the GEMM's copy-engine loader on FULL_B[1] and traps -- in a process of
    its own, because the CUDA context is unusable afterwards."""
    import subprocess
    import sys
    code = r"""
import ctypes, sys
sys.path.insert(0, %r)
import torch
from ophelia_b200 import _lib
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device="cuda")
rc = lib.oph_debug_trap_selftest(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
raw = (ctypes.c_ulonglong * 4)()
text = ctypes.create_string_buffer(512)
have = lib.oph_last_trap(raw, text, 512)
print("RC", rc, "HAVE", have)
print("ERR", lib.oph_last_error().decode())
print("TEXT", text.value.decode())
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    assert "HAVE 1" in out and "RC 0" not in out, out
    assert "copy-engine loader" in out and "FULL_B[1]" in out and "parity 1" in out, out
    assert "ERR trap_selftest_kernel" in out and "bounded barrier wait trapped" in out.split("ERR", 1)[1], out
