"""Oracle: mono mix-down + sinc resampling.  TEST INFRASTRUCTURE ONLY.

The reference converts every decoded file with `convert_audio` (audiotoken/utils.py:26-44): stereo -> mean over
channels, then `torchaudio.transforms.Resample(sample_rate, target)` with torchaudio's defaults
(resampling_method='sinc_interp_hann', lowpass_filter_width=6, rolloff=0.99); the streaming reader uses the same
transform per chunk (utils.py:98-99).  torchaudio is a third-party dependency (requirements.txt); its published
algorithm (torchaudio/functional/functional.py::_get_sinc_resample_kernel / _apply_sinc_resample_kernel) is restated
here in numpy: the rates are reduced by their gcd, one windowed-sinc filter per output phase is built in float64 and
rounded to float32, the input is zero-padded (width, width + orig) and correlated with stride orig.
tests/golden/make_golden_resample.py stores outputs of torchaudio itself; tests/test_oracle_golden.py pins this file
against them.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99


def reduced_rates(orig_freq: int, new_freq: int) -> Tuple[int, int]:
    g = math.gcd(int(orig_freq), int(new_freq))
    return int(orig_freq) // g, int(new_freq) // g


def sinc_kernel(orig_freq: int, new_freq: int) -> Tuple[np.ndarray, int]:
    """-> (kernels float32 [new, 2*width + orig], width) for gcd-reduced rates."""
    orig, new = reduced_rates(orig_freq, new_freq)
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = t * base_freq
    t = np.clip(t, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    safe = np.where(t == 0, 1.0, t)
    kernels = np.where(t == 0, 1.0, np.sin(safe) / safe) * window * scale
    return kernels.astype(np.float32), width


def convert_audio(audio: np.ndarray, sample_rate: int, target_sample_rate: int) -> np.ndarray:
    """audio float32 [C, L] (C = 1 or 2) -> float32 [1, ceil(new * L / orig)]  (reference utils.py:26-44)."""
    audio = np.asarray(audio, dtype=np.float32)
    c, length = audio.shape
    if c == 2:
        audio = ((audio[0] + audio[1]) / np.float32(2.0))[None, :]
    elif c != 1:
        raise RuntimeError('Only mono or stereo audio is supported')
    if sample_rate == target_sample_rate:
        return audio
    orig, new = reduced_rates(sample_rate, target_sample_rate)
    kernels, width = sinc_kernel(sample_rate, target_sample_rate)
    x = np.zeros(width + length + width + orig, dtype=np.float32)
    x[width:width + length] = audio[0]
    n_frames = (x.size - kernels.shape[1]) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(x, kernels.shape[1])[::orig][:n_frames]      # [frames, K]
    out = (win.astype(np.float32) @ kernels.T.astype(np.float32)).reshape(-1)                   # frame-major, phase-minor
    target_length = int(math.ceil(new * length / orig))
    return out[:target_length][None, :].astype(np.float32)
