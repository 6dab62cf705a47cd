"""Griffin-Lim vocoder of the reference (`utils.py:69-116`, called from `synthesize.py:433-440`) on the GPU
(SURVEY 8(f) next-4).

The reference runs 50 iterations of librosa istft/stft per utterance on CPU worker processes.  Here the whole iteration
stays on the device: `torch.stft` / `torch.istft` (cuFFT) follow librosa's conventions -- centred frames with reflect
padding, periodic Hann window of `win_length` zero-padded to `n_fft`, inverse by windowed overlap-add divided by the
summed squared window, `n_fft // 2` samples trimmed from both ends.  librosa is not in this image; its algorithm is
restated with numpy FFTs in `oracle/griffin_lim_numpy.py`, and tests compare the two.

The FFTs are library calls (cuFFT through torch): this stage comes after the hot path and is not one of its kernels.
"""
import numpy as np
import torch


def _window(hp, device, dtype=torch.float32):
    return torch.hann_window(hp.win_length, periodic=True, device=device, dtype=dtype)


def stft(hp, y, window=None):
    """librosa.stft(y, n_fft, hop_length, win_length): complex [1 + n_fft // 2, frames]."""
    window = _window(hp, y.device, y.dtype) if window is None else window
    return torch.stft(y, hp.n_fft, hop_length=hp.hop_length, win_length=hp.win_length, window=window, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)


def invert_spectrogram(hp, spectrogram, window=None):
    """utils.py:111-116: librosa.istft(spectrogram, hop_length, win_length, window='hann')."""
    window = _window(hp, spectrogram.device, spectrogram.real.dtype) if window is None else window
    return torch.istft(spectrogram, hp.n_fft, hop_length=hp.hop_length, win_length=hp.win_length, window=window,
                       center=True, normalized=False, onesided=True, return_complex=False)


def griffin_lim(hp, spectrogram):
    """utils.py:98-109.  spectrogram: real magnitudes [1 + n_fft // 2, frames] on any device, fp32 (fp64 is accepted:
    the phase of a near-silent bin is ill-conditioned, so bit-level comparisons with the fp64 restatement need it);
    returns the waveform (same device).  Every iteration re-imposes the given magnitudes on the phase of the re-analysed
    signal."""
    window = _window(hp, spectrogram.device, spectrogram.dtype)
    X_best = spectrogram.to(torch.complex128 if spectrogram.dtype == torch.float64 else torch.complex64)
    for _ in range(hp.n_iter):
        X_t = invert_spectrogram(hp, X_best, window)
        est = stft(hp, X_t, window)
        phase = est / torch.clamp(est.abs(), min=1e-8)
        X_best = spectrogram * phase
    return invert_spectrogram(hp, X_best, window)


def deemphasis(wav, preemphasis):
    """signal.lfilter([1], [1, -preemphasis], wav) (utils.py:91): y[n] = x[n] + p * y[n - 1]."""
    from scipy import signal
    return signal.lfilter([1], [1, -preemphasis], wav)


def spectrogram2wav(hp, mag, trim_output=False, device=None, dtype=torch.float32):
    """utils.py:69-96.  mag: normalised magnitudes [frames, 1 + n_fft // 2] in [0, 1] (SSRN output); returns float32
    samples.  Runs on the GPU; there is no silent CPU fallback (tests pass device='cpu' explicitly to compare the same
    code with the numpy restatement)."""
    assert not trim_output, "librosa.effects.trim is outside the path (the generation loop already stops at the sentence end)"
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("spectrogram2wav: no CUDA device (pass device='cpu' explicitly for an offline check)")
        device = torch.device("cuda")
    mag = torch.as_tensor(np.ascontiguousarray(mag), dtype=dtype).to(device).t()
    mag = torch.clamp(mag, 0, 1) * hp.max_db - hp.max_db + hp.ref_db          # de-normalise (dB)
    mag = torch.pow(10.0, mag * 0.05)                                          # to amplitude
    wav = griffin_lim(hp, mag ** hp.power)
    wav = deemphasis(wav.cpu().numpy().astype(np.float64), hp.preemphasis)
    return wav.astype(np.float32)


def write_wav(path, wav, sr):
    """soundfile.write(path, wav, sr) of synthesize.py:438 (16-bit PCM, the subtype soundfile picks for .wav)."""
    from scipy.io import wavfile
    pcm = np.clip(np.asarray(wav, np.float64), -1.0, 1.0 - 1.0 / 32768.0)
    wavfile.write(path, int(sr), np.round(pcm * 32768.0).astype(np.int16))


def synth_wave(hp, mag, outfile, device=None):
    """synthesize.py:433-440 for hp.vocoder == 'griffin_lim' (the WORLD vocoder is an external binary: outside the path)."""
    assert hp.vocoder == 'griffin_lim', "only the Griffin-Lim vocoder is inside the path"
    wav = spectrogram2wav(hp, mag, device=device)
    if getattr(hp, "store_synth_features", False):
        np.save(outfile.replace('.wav', '.npy'), mag)
    write_wav(outfile, wav, hp.sr)
    return wav
