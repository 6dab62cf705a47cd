"""Text2Mel forward + one training step against the oracles over shapes that sit on tile (128 rows), k-block (64) and
batch-item boundaries, including the degenerate ones (one frame, a handful of phonemes): the reference pads every batch
to its own (N_b, T_b) (data_load.py:534-541), so any shape can arrive."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 5, 1), (1, 9, 127), (2, 64, 128), (3, 65, 129), (2, 33, 255), (5, 17, 64)])
def test_awkward_shapes_match_oracle(shape):
    import shape_fuzz
    eY, eA, eL = shape_fuzz.check(*shape)
    assert eY < 1e-3 and eA < 1e-4 and eL < 5e-4, (shape, eY, eA, eL)
