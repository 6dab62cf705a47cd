"""Griffin-Lim vocoder (utils.py:69-116): the torch/cuFFT implementation against the numpy restatement of librosa's
stft / istft, on CPU here and on the GPU under -m gpu."""
import numpy as np
import pytest
import torch

from ophelia_b200 import vocoder
from ophelia_b200.configuration import default_hparams
from oracle import griffin_lim_numpy as gl


def _hp(**kw):
    base = dict(n_fft=256, hop_length=32, win_length=128, n_iter=8, power=1.5, preemphasis=0.97, max_db=100, ref_db=20,
                vocoder='griffin_lim', sr=8000)
    base.update(kw)
    return default_hparams(**base)


def test_numpy_restatement_reconstructs_the_signal():
    """Known answer of the stft / istft pair: with a window that satisfies the overlap-add condition the inverse of the
    forward transform is the signal itself (up to the last partial hop)."""
    rng = np.random.default_rng(0)
    y = rng.normal(size=2000)
    S = gl.stft(y, 256, 32, 128)
    assert S.shape == (129, 1 + len(y) // 32)
    back = gl.istft(S, 32, 128)
    assert len(back) == 32 * (S.shape[1] - 1)
    np.testing.assert_allclose(back, y[:len(back)], atol=1e-10)


def _check(device, tol):
    hp = _hp()
    rng = np.random.default_rng(1)
    y = rng.normal(size=1500)
    S_ref = gl.stft(y, hp.n_fft, hp.hop_length, hp.win_length)
    S = vocoder.stft(hp, torch.tensor(y, dtype=torch.float32, device=device)).cpu().numpy()
    assert S.shape == S_ref.shape and np.abs(S - S_ref).max() < 1e-3 * tol
    back = vocoder.invert_spectrogram(hp, torch.tensor(S_ref, dtype=torch.complex64, device=device)).cpu().numpy()
    np.testing.assert_allclose(back, gl.istft(S_ref, hp.hop_length, hp.win_length), atol=1e-4 * tol)
    # the full chain on an SSRN-like magnitude track: de-normalise, amplify, 8 Griffin-Lim iterations, de-emphasis.
    # The phase of a near-silent bin is ill-conditioned (est / max(1e-8, |est|)), so the fp32 product path cannot be
    # compared sample by sample with the fp64 restatement: the algorithm is checked in fp64 (exact), the fp32 path
    # through the spectrum it converges to
    t = np.arange(40)[:, None]
    f = np.arange(129)[None, :]
    mag = np.clip(0.55 + 0.3 * np.sin(0.3 * t + 0.05 * f) * np.exp(-f / 90.0), 0, 1).astype(np.float32)
    w8_ref = gl.spectrogram2wav(hp, mag)
    w64 = vocoder.spectrogram2wav(hp, mag, device=device, dtype=torch.float64)
    assert w64.shape == w8_ref.shape == (hp.hop_length * (len(mag) - 1),)
    assert np.abs(w64 - w8_ref).max() < 1e-6 * np.abs(w8_ref).max()
    w8 = vocoder.spectrogram2wav(hp, mag, device=device)
    m8 = np.abs(gl.stft(w8, hp.n_fft, hp.hop_length, hp.win_length))
    m8_ref = np.abs(gl.stft(w8_ref, hp.n_fft, hp.hop_length, hp.win_length))
    assert np.linalg.norm(m8 - m8_ref) / np.linalg.norm(m8_ref) < 0.05
    return w8


def test_torch_vocoder_matches_numpy_restatement_cpu(tmp_path):
    w = _check(torch.device("cpu"), 1.0)
    path = str(tmp_path / "x.wav")
    vocoder.write_wav(path, w / max(1.0, np.abs(w).max()), 8000)
    from scipy.io import wavfile
    sr, pcm = wavfile.read(path)
    assert sr == 8000 and pcm.dtype == np.int16 and len(pcm) == len(w)


@pytest.mark.gpu
def test_torch_vocoder_matches_numpy_restatement_gpu():
    _check(torch.device("cuda"), 2.0)
