"""Device-side audio ingest: PCM decode + mono mix-down + resampling (reference audiotoken/utils.py:26-44, 98-99).

`GpuResampler(orig, new)` builds torchaudio's `sinc_interp_hann` phase filters (lowpass_filter_width 6, rolloff 0.99)
on the host in float64 -> float32, strips them to their non-zero support and keeps them on the device;
`convert_audio` is the drop-in for the reference helper of the same name with the arithmetic in libb200tok.so
(csrc/ingest.cu).  There is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import lib as L

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99


def _phase_filters(orig: int, new: int) -> Tuple[np.ndarray, int]:
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx) * base_freq
    t = np.clip(t, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    safe = np.where(t == 0, 1.0, t)
    k = np.where(t == 0, 1.0, np.sin(safe) / safe) * window * (base_freq / orig)
    return k.astype(np.float32), width


class GpuResampler:
    def __init__(self, orig_freq: int, new_freq: int, device='cuda:0'):
        self.device = torch.device(device)
        L.require_device(self.device)
        g = math.gcd(int(orig_freq), int(new_freq))
        self.orig, self.new = int(orig_freq) // g, int(new_freq) // g
        if self.orig == self.new:                      # the reference skips Resample for equal rates (utils.py:41)
            k, self.width = np.ones((1, 1), dtype=np.float32), 0
        else:
            k, self.width = _phase_filters(self.orig, self.new)
        nz = k != 0
        start = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
        last = np.where(nz.any(1), k.shape[1] - 1 - nz[:, ::-1].argmax(1), 0)
        count = np.where(nz.any(1), last - start + 1, 0).astype(np.int32)
        self.max_taps = int(max(1, count.max()))
        taps = np.zeros((self.new, self.max_taps), dtype=np.float32)
        for i in range(self.new):
            taps[i, :count[i]] = k[i, start[i]:start[i] + count[i]]
        self.taps = torch.from_numpy(taps).to(self.device)
        self.start = torch.from_numpy(start).to(self.device)
        self.count = torch.from_numpy(count).to(self.device)

    def out_len(self, length: int) -> int:
        return int(math.ceil(self.new * length / self.orig))

    def __call__(self, audio: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio [C, L] (any strides, e.g. the transposed view of an interleaved PCM buffer), int16 or float32, on
        the device -> float32 [1, L'] on the device (`out`: caller-owned contiguous [1, L'] destination, e.g. a slice of
        one buffer per window of files — the streaming loop must not allocate per file, see core.encode_files)."""
        assert audio.is_cuda and audio.dim() == 2 and audio.dtype in (torch.int16, torch.float32)
        c, length = audio.shape
        if c not in (1, 2):
            raise RuntimeError('Only mono or stereo audio is supported')
        if out is None:
            out = torch.empty(1, self.out_len(length), dtype=torch.float32, device=audio.device)
        elif tuple(out.shape) != (1, self.out_len(length)) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f'out must be a contiguous float32 [1, {self.out_len(length)}] tensor')
        if out.numel() == 0:
            return out
        with torch.cuda.device(audio.device):
            L.check(L.load().b2t_ingest_resample(audio.data_ptr(), int(audio.dtype == torch.int16), length, c,
                                                 audio.stride(0) if c == 2 else 0, audio.stride(1), self.taps.data_ptr(),
                                                 self.start.data_ptr(), self.count.data_ptr(), self.max_taps, self.orig,
                                                 self.new, self.width, out.data_ptr(), out.shape[1], L.stream_ptr()),
                    'b2t_ingest_resample')
        return out


_CACHE: Dict[Tuple[int, int, str], GpuResampler] = {}


def resampler(sample_rate: int, target_sample_rate: int, device='cuda:0') -> GpuResampler:
    dev = torch.device(device)
    key = (int(sample_rate), int(target_sample_rate), str(dev))
    if key not in _CACHE:
        _CACHE[key] = GpuResampler(sample_rate, target_sample_rate, dev)
    return _CACHE[key]


def convert_audio(audio: torch.Tensor, sample_rate: int, target_sample_rate: int, device='cuda:0',
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[C, L] int16 PCM or float32 -> mono float32 [1, L'] at the target rate, on `device` (reference utils.py:26-44).
    Equal rates run the same kernel with the identity filter (decode + mix-down only)."""
    dev = torch.device(device)
    if audio.shape[0] not in (1, 2):
        raise RuntimeError('Only mono or stereo audio is supported')
    return resampler(sample_rate, target_sample_rate, dev)(audio.to(dev, non_blocking=True), out)
