"""Synthetic batches that reproduce the output contract of the reference's input pipeline
(`data_load.py:302-544`: text int32 [B,N] zero-padded, mel fp32 [B,T,n_mels] in [1e-8,1] zero-padded,
mag fp32 [B,T*r,full_dim]) -- what bench.py and the parity tests feed.  The pipeline over real transcripts and .npy
features is ophelia_b200/data_load.py."""
import numpy as np
import torch


class SyntheticBatches(object):
    """Endless iterator over `n_distinct` pre-generated host batches in pinned memory."""

    def __init__(self, hp, model, batch, N=180, T=870, seed=1234, n_distinct=2, num_batch=400, min_text=120,
                 ragged=True, pin=True):
        rng = np.random.default_rng(seed)
        V = len(hp.vocab)
        self.num_batch = num_batch
        self.batches = []
        for _ in range(n_distinct):
            b = {}
            if model == "t2m":
                L = np.zeros((batch, N), np.int32)
                for i in range(batch):
                    n = int(rng.integers(min(min_text, N), N + 1))
                    L[i, :n] = rng.integers(1, V, n)
                b["text"] = L
            mel = rng.uniform(1e-8, 1.0, (batch, T, hp.n_mels)).astype(np.float32)
            if ragged:
                for i in range(batch):
                    mel[i, int(rng.integers((2 * T) // 3, T + 1)):] = 0.0
            b["mel"] = mel
            if model == "ssrn":
                b["mag"] = rng.uniform(1e-8, 1.0, (batch, T * hp.r, hp.full_dim)).astype(np.float32)
            tb = {k: torch.from_numpy(v) for k, v in b.items()}
            if pin and torch.cuda.is_available():
                tb = {k: v.pin_memory() for k, v in tb.items()}
            self.batches.append(tb)
        self._i = 0

    def __iter__(self):
        return self

    def __next__(self):
        b = self.batches[self._i % len(self.batches)]
        self._i += 1
        return b

    def bytes_per_batch(self):
        return sum(v.numel() * v.element_size() for v in self.batches[0].values())
