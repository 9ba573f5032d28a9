"""Constant tables of the log-mel front end, in the form the CUDA kernel consumes.

Same definitions as the reference (audiotoken/processors.py:8-26, 66-78 and the Kaldi mel helpers
audiotoken/utils.py:286-328): povey window = hann(400, symmetric)^0.85; 80 triangular filters whose
82 edge points are linearly spaced in *mel* between mel(20 Hz) and mel(8000 Hz), evaluated at the
mel value of the 256 FFT-bin centres (31.25 Hz apart); the Nyquist bin has weight 0.  The dense
[257, 80] bank is converted to (first bin, count, weights) per filter — each filter touches at most
32 consecutive bins (16 in practice).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

N_MEL, MAX_SPAN, N_BINS = 80, 32, 256


def povey_window() -> torch.Tensor:
    return torch.pow(torch.hann_window(400, periodic=False, dtype=torch.float32), 0.85)


def _mel(f: torch.Tensor) -> torch.Tensor:
    return 1127.0 * torch.log(1.0 + f / 700.0)


def dense_mel_bank(sample_rate: int = 16000) -> torch.Tensor:
    """[256, 80] fp32 (the Nyquist row of the reference's [257, 80] matrix is identically 0)."""
    edges = torch.linspace(_mel(torch.tensor(20.0)), _mel(torch.tensor(float(sample_rate // 2))), N_MEL + 2)
    centres = _mel((sample_rate / (2 * N_BINS)) * torch.arange(N_BINS, dtype=torch.float32))
    width = edges[1:] - edges[:-1]
    delta = edges.unsqueeze(0) - centres.unsqueeze(1)                 # [256, 82]
    falling = -delta[:, :-2] / width[:-1]
    rising = delta[:, 2:] / width[1:]
    return torch.clamp(torch.minimum(falling, rising), min=0.0)


def sparse_mel_bank(dense: torch.Tensor):
    """dense [bins, 80] -> (start int32[80], count int32[80], weight fp32[80, 32])."""
    d = dense.numpy()
    start = np.zeros(N_MEL, dtype=np.int32)
    count = np.zeros(N_MEL, dtype=np.int32)
    weight = np.zeros((N_MEL, MAX_SPAN), dtype=np.float32)
    for f in range(N_MEL):
        nz = np.nonzero(d[:, f])[0]
        if nz.size == 0:
            continue
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        if hi - lo > MAX_SPAN:
            raise ValueError(f'mel filter {f} spans {hi - lo} bins (> {MAX_SPAN})')
        start[f], count[f] = lo, hi - lo
        weight[f, :hi - lo] = d[lo:hi, f]
    return start, count, weight


class DeviceFbankTables:
    def __init__(self, device, sample_rate: int = 16000):
        from . import lib as L
        start, count, weight = sparse_mel_bank(dense_mel_bank(sample_rate))
        self.window = povey_window().to(device)
        self.start = torch.from_numpy(start).to(device)
        self.count = torch.from_numpy(count).to(device)
        self.weight = torch.from_numpy(weight).to(device)
        t = L.FbankTables()
        t.window, t.mel_start, t.mel_count, t.mel_weight = (
            self.window.data_ptr(), self.start.data_ptr(), self.count.data_ptr(), self.weight.data_ptr())
        self.c = t

    def byref(self):
        return C.byref(self.c)
