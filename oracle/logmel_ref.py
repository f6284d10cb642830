"""float64 ground truth for Whisper log-mel features.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates HF/models/whisper/feature_extraction_whisper.py:135-164 (`_torch_extract_fbank_features`; its numpy twin
`:105-133` is the same arithmetic in float64) and the Slaney filterbank of HF/audio_utils.py:285-332,453-544.
All arithmetic up to the final cast is float64, so this is the "truth" against which both HF paths and the
CUDA kernel are measured (SURVEY.md §7 hard part 2: HF-numpy is within 1.2e-7 of it, HF-torch within 3.7e-5).
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000
N_FFT = 400
HOP = 160
N_FREQ = N_FFT // 2 + 1


def _hz_to_mel(f):
    # HF/audio_utils.py:285-296 (slaney): linear below 1 kHz, log above
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    with np.errstate(divide="ignore", invalid="ignore"):
        log = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    # HF/audio_utils.py:321-332
    m = np.asarray(m, dtype=np.float64)
    lin = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    log = 1000.0 * np.exp(logstep * (m - 15.0))
    return np.where(m >= 15.0, log, lin)


def mel_filter_bank(n_mels: int, n_freq: int = N_FREQ, fmin: float = 0.0, fmax: float = 8000.0,
                    sampling_rate: int = SAMPLING_RATE) -> np.ndarray:
    """[n_freq, n_mels] float64 triangular Slaney-normalised filters (HF/audio_utils.py:453-544)."""
    mel_pts = np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2)
    filter_freqs = _mel_to_hz(mel_pts)
    fft_freqs = np.linspace(0.0, sampling_rate // 2, n_freq)
    fdiff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    return fb * enorm[None, :]


def pad_or_trim(audio: np.ndarray, n_samples: int) -> np.ndarray:
    """float32 clip right-padded with zeros / truncated to n_samples (feature_extraction_whisper.py:296-303)."""
    a = np.asarray(audio, dtype=np.float32).reshape(-1)
    if a.shape[0] >= n_samples:
        return a[:n_samples].copy()
    out = np.zeros(n_samples, dtype=np.float32)
    out[: a.shape[0]] = a
    return out


def power_spectrogram_f64(x: np.ndarray) -> np.ndarray:
    """|STFT|^2 in float64: [201, L//160] (centre reflect pad, periodic Hann, last frame dropped)."""
    x = np.asarray(x, dtype=np.float64)
    L = x.shape[0]
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + L // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    n = np.arange(N_FFT, dtype=np.float64)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)  # periodic Hann
    spec = np.fft.rfft(xp[idx] * window[None, :], axis=1)  # [frames, 201]
    power = spec.real ** 2 + spec.imag ** 2
    return power[:-1].T  # drop frame L//160 (feature_extraction_whisper.py:150)


def logmel_f64(audio: np.ndarray, n_mels: int, n_samples: int = 480000) -> np.ndarray:
    """One clip -> float32 [n_mels, n_samples//160]; everything before the final cast in float64."""
    x = pad_or_trim(audio, n_samples)
    power = power_spectrogram_f64(x)
    mel = mel_filter_bank(n_mels).T @ power
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)  # per-clip max (feature_extraction_whisper.py:156-160)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def logmel_batch_f64(clips, n_mels: int, n_samples: int = 480000) -> np.ndarray:
    return np.stack([logmel_f64(c, n_mels, n_samples) for c in clips], axis=0)


def frame_attention_mask(lengths, n_samples: int = 480000) -> np.ndarray:
    """int32 [B, n_samples//160]: sample mask subsampled every hop (feature_extraction_whisper.py:328-337)."""
    lengths = np.minimum(np.asarray(lengths, dtype=np.int64), n_samples)
    m = (np.arange(n_samples)[None, :] < lengths[:, None]).astype(np.int32)[:, ::HOP]
    if n_samples % HOP != 0:
        m = m[:, :-1]
    return m
