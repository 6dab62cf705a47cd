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
