"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy restatement of the reference's Griffin-Lim vocoder
(`utils.py:69-116`), whose arithmetic lives in a dependency absent from /root/reference and from this image:
librosa (requirements.txt: librosa==0.6.2).  Restated from librosa 0.6's published algorithm:

* stft: y reflect-padded by n_fft // 2 on both sides; frame t = y_pad[t * hop : t * hop + n_fft] * w, where w is the periodic
  ('fftbins') Hann window of win_length samples zero-padded symmetrically to n_fft; rfft of every frame.
* istft: irfft of every column times w, overlap-added at t * hop into a buffer of n_fft + hop * (frames - 1) samples,
  divided by the overlap-added w^2 wherever that exceeds `tiny(float32)`, then n_fft // 2 samples cut from both ends.
Parity unpinned against librosa itself (not installable here); the known answer it must satisfy is perfect
reconstruction istft(stft(y)) == y away from the edges, checked in tests/test_vocoder.py.
"""
import numpy as np


def hann_padded(win_length, n_fft):
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)      # scipy.signal.get_window('hann', fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad), mode="constant")          # librosa.util.pad_center


def stft(y, n_fft, hop_length, win_length):
    w = hann_padded(win_length, n_fft)
    yp = np.pad(np.asarray(y, np.float64), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    out = np.empty((1 + n_fft // 2, n_frames), np.complex128)
    for t in range(n_frames):
        out[:, t] = np.fft.rfft(yp[t * hop_length:t * hop_length + n_fft] * w)
    return out


def istft(S, hop_length, win_length):
    n_fft = 2 * (S.shape[0] - 1)
    w = hann_padded(win_length, n_fft)
    n_frames = S.shape[1]
    y = np.zeros(n_fft + hop_length * (n_frames - 1))
    wss = np.zeros_like(y)
    for t in range(n_frames):
        y[t * hop_length:t * hop_length + n_fft] += w * np.fft.irfft(S[:, t], n_fft)
        wss[t * hop_length:t * hop_length + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:len(y) - n_fft // 2]


def griffin_lim(hp, spectrogram):
    """utils.py:98-109."""
    spectrogram = np.asarray(spectrogram, np.float64)
    X_best = spectrogram.astype(np.complex128)
    for _ in range(hp.n_iter):
        X_t = istft(X_best, hp.hop_length, hp.win_length)
        est = stft(X_t, hp.n_fft, hp.hop_length, hp.win_length)
        phase = est / np.maximum(1e-8, np.abs(est))
        X_best = spectrogram * phase
    return np.real(istft(X_best, hp.hop_length, hp.win_length))


def spectrogram2wav(hp, mag):
    """utils.py:69-96 (trim_output=False)."""
    mag = np.asarray(mag, np.float64).T
    mag = (np.clip(mag, 0, 1) * hp.max_db) - hp.max_db + hp.ref_db
    mag = np.power(10.0, mag * 0.05)
    wav = griffin_lim(hp, mag ** hp.power)
    out = np.empty_like(wav)                      # scipy.signal.lfilter([1], [1, -preemphasis], wav)
    prev = 0.0
    for i, x in enumerate(wav):
        prev = x + hp.preemphasis * prev
        out[i] = prev
    return out.astype(np.float32)
