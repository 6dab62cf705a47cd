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
