"""Slaney mel filterbank used by WhisperFeatureExtractor (restated; no transformers import).

Follows HF/audio_utils.py:453-544 `mel_filter_bank(201, n_mels, 0, 8000, 16000, norm="slaney", mel_scale="slaney")`
as called from HF/models/whisper/feature_extraction_whisper.py:84-92. Pinned against the HF function in
tests/test_oracle_cpu.py (known answers in SURVEY.md 8(c): 391 non-zeros, max 0.025880684545274913).
"""
from __future__ import annotations

import numpy as np

_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = 15.0
_LOGSTEP = 27.0 / np.log(6.4)


def _hz_to_mel(f: float) -> float:
    return _MIN_LOG_MEL + np.log(f / _MIN_LOG_HZ) * _LOGSTEP if f >= _MIN_LOG_HZ else 3.0 * f / 200.0


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    hz = 200.0 * m / 3.0
    log_region = m >= _MIN_LOG_MEL
    hz[log_region] = _MIN_LOG_HZ * np.exp((m[log_region] - _MIN_LOG_MEL) / _LOGSTEP)
    return hz


def whisper_mel_filters(n_mels: int = 80, n_bins: int = 201, sr: int = 16000, fmax: float = 8000.0) -> np.ndarray:
    """float64 [n_bins, n_mels] triangular, area-normalised filters (cast to float32 by the consumer)."""
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(fmax), n_mels + 2)
    edges = _mel_to_hz(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_bins)
    diff = np.diff(edges)
    slopes = edges[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[None, :]
    return fb
