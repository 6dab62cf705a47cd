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
