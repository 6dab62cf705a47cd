"""Generates the committed golden vectors under tests/golden/ from the numpy fp64 oracle.

PARITY UNPINNED: these pin the CUDA path to OUR restatement of the reference (no TF-1.12 output exists to pin the
restatement itself, see oracle/__init__.py).  Run from the repo root:  python -m oracle.make_golden
"""
import os

import numpy as np

from oracle import dctts_numpy as on
from oracle.params import HP, init_params, ssrn_specs, synthetic_batch, text2mel_specs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # Text2Mel, BASELINE config 1: batch 2, 60 phonemes (50 real), 200 mel frames
    hp = HP(max_N=60, max_T=200)
    P = init_params(text2mel_specs(hp), 7, perturb=True)
    b = synthetic_batch(hp, 2, 60, 200, seed=4321, text_len=50)
    out = on.text2mel_forward(hp, P, b["L"], b["mels"], "generate_attention")
    comps = on.text2mel_loss(hp, out, b["mels"])
    np.savez_compressed(os.path.join(OUT, "t2m_c1.npz"), param_seed=7, data_seed=4321,
                        Y=out["Y"].astype(np.float32), alignments=out["alignments"].astype(np.float32),
                        max_attentions=out["max_attentions"].astype(np.int32),
                        K_sum=np.float64(out["K"].sum()), Q_sum=np.float64(out["Q"].sum()),
                        loss_components=np.asarray(comps, np.float64))
    # SSRN: batch 2, 24 -> 96 frames, 513 bins
    hp2 = HP(full_dim=513)
    Ps = init_params(ssrn_specs(hp2), 8, perturb=True)
    b2 = synthetic_batch(hp2, 2, 8, 24, seed=99, with_mags=True)
    logits, Z = on.SSRN(hp2, Ps, b2["mels"].astype(np.float64))
    comps2 = on.ssrn_loss(hp2, logits, Z, b2["mags"])
    np.savez_compressed(os.path.join(OUT, "ssrn_small.npz"), param_seed=8, data_seed=99,
                        Z=Z.astype(np.float32), loss_components=np.asarray(comps2, np.float64))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
