"""Oracle: SeamlessM4T/Kaldi-style log-mel front end.  TEST INFRASTRUCTURE ONLY.

Restates ``audiotoken/processors.py`` (reference):
  * mel filter bank ............ processors.py:8-26 + utils.py:286-328
  * window / buffers ........... processors.py:66-78
  * frame loop ................. processors.py:155-178 (vectorised here: same per-frame maths)
  * power / mel / floor / log .. processors.py:181-188
  * frame validity mask ........ processors.py:80-115
  * masked mean / biased var ... processors.py:117-135, 241-242
  * stride-2 stacking + pad .... processors.py:244-259, 192-207
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

FRAME = 400
HOP = 160
NFFT = 512
NBINS = 257
NMEL = 80
PREEMPH = 0.97
MEL_FLOOR = 1.192092955078125e-07
SR = 16000


def hz_to_mel(f: torch.Tensor) -> torch.Tensor:
    # utils.py:286-296 (Kaldi scale)
    return 1127.0 * torch.log(1.0 + (f / 700.0))


def mel_filters(dtype=torch.float32) -> torch.Tensor:
    """[257, 80] triangles built in mel space; the Nyquist row is zero.

    processors.py:16-26: 82 points linearly spaced *in mel* between mel(20) and mel(8000)
    (the Hz conversion at :19 is overwritten at :21), FFT-bin centres mel(31.25*k), k<256;
    utils.py:323-328: tri = max(0, min(down, up)); processors.py:77 appends the zero row.
    """
    mel_min = hz_to_mel(torch.tensor(20.0, dtype=dtype))
    mel_max = hz_to_mel(torch.tensor(float(SR // 2), dtype=dtype))
    filt = torch.linspace(mel_min, mel_max, NMEL + 2, dtype=dtype)
    bin_width = SR / (256 * 2)
    fft_freqs = hz_to_mel(bin_width * torch.arange(256, dtype=dtype))
    diff = torch.diff(filt)
    slopes = filt.unsqueeze(0) - fft_freqs.unsqueeze(1)          # [256, 82]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    tri = torch.maximum(torch.zeros(1, dtype=dtype), torch.minimum(down, up))
    return torch.cat([tri, torch.zeros(1, NMEL, dtype=dtype)], dim=0)


def povey_window(dtype=torch.float32) -> torch.Tensor:
    # processors.py:75: hann(400, periodic=False) ** 0.85
    return torch.pow(torch.hann_window(FRAME, periodic=False, dtype=dtype), 0.85)


def num_frames(num_samples: int) -> int:
    # processors.py:158
    return int(1 + math.floor((num_samples - FRAME) / HOP)) if num_samples >= FRAME else 0


def log_mel(wave: torch.Tensor, dtype=torch.float32, mel_in_bf16: bool = False) -> torch.Tensor:
    """wave [B, L] in [-1, 1] -> log-mel [B, N, 80]  (processors.py:137-190).

    `mel_in_bf16` reproduces the one autocast cast point of the front end on CUDA: the
    ``matmul(power, mel_filters)`` at processors.py:184 runs with bf16 inputs/outputs.
    Device-generic: on a CUDA `wave` under ``torch.amp.autocast`` the plain matmul below is cast by autocast itself
    (oracle/hf_reference.py runs it that way, as the reference is deployed).
    """
    dev = wave.device
    w = wave.to(dtype) * (2 ** 15)
    B, L = w.shape
    N = num_frames(L)
    fr = w.unfold(1, FRAME, HOP)[:, :N].clone()                  # [B, N, 400]
    fr = fr - fr.mean(dim=2, keepdim=True)                       # :168-169
    pre = fr.clone()
    pre[:, :, 1:] = fr[:, :, 1:] - PREEMPH * fr[:, :, :-1]       # :171-172 (RHS = old values)
    pre[:, :, 0] = fr[:, :, 0] * (1 - PREEMPH)                   # :173
    pre = pre * povey_window(dtype).to(dev)                      # :175
    buf = torch.zeros(B, N, NFFT, dtype=dtype, device=dev)
    buf[:, :, :FRAME] = pre
    spec = torch.fft.rfft(buf)                                   # :177
    power = spec.abs().pow(2.0)                                  # :181
    M = mel_filters(dtype).to(dev)
    if mel_in_bf16:
        mel = (power.to(torch.bfloat16).float() @ M.to(torch.bfloat16).float())
        mel = mel.to(torch.bfloat16).to(dtype)
    else:
        mel = power @ M                                          # :184
    mel = torch.maximum(mel.to(dtype), torch.tensor(MEL_FLOOR, dtype=dtype, device=dev))
    return torch.log(mel)                                        # :188


def frame_mask(mask: torch.Tensor, n_frames: int) -> torch.Tensor:
    """Sample mask [B, L] (0/1) -> frame validity [B, N]: valid iff all 400 samples valid
    (processors.py:102-108: avg_pool1d(k=400, s=160) == 1)."""
    m = mask.float().unfold(1, FRAME, HOP)[:, :n_frames]
    return (m.mean(dim=2) == 1).float()


def features(wave: torch.Tensor, mask: torch.Tensor, pad_to_multiple_of: int = 2,
             dtype=torch.float32, mel_in_bf16: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Full processor: (input_features [B, T, 160], attention_mask [B, T]).

    processors.py:209-266.
    """
    x = log_mel(wave, dtype, mel_in_bf16)                        # [B, N, 80]
    B, N, C = x.shape
    fm = frame_mask(mask, N).to(dtype).unsqueeze(-1).expand(-1, -1, C)
    # masked mean / biased variance over valid frames (:117-135)
    mx = x * fm
    cnt = fm.sum(dim=1, keepdim=True).clamp(min=1)
    mean = mx.sum(dim=1, keepdim=True) / cnt
    var = (((mx - mean) ** 2) * fm).sum(dim=1, keepdim=True) / cnt
    x = (x - mean) / torch.sqrt(var + 1e-7)                      # :242 (all frames)
    rem = N % 2                                                  # :244-250
    if rem:
        x = x[:, :N - rem]
        fm = fm[:, :N - rem]
    x = x.reshape(B, (N - rem) // 2, 2 * C)                      # :252-257
    fm = fm.reshape(B, (N - rem) // 2, 2 * C)
    T = x.shape[1]
    P = 0
    if pad_to_multiple_of > 0 and T % pad_to_multiple_of:
        P = pad_to_multiple_of - T % pad_to_multiple_of
    x = torch.where(fm == 0, torch.tensor(1.0, dtype=dtype, device=x.device), x)  # :200 padding_value = 1
    x = torch.nn.functional.pad(x, (0, 0, 0, P), value=1.0)      # :201
    am = torch.nn.functional.pad(fm[:, :, 0], (0, P), value=0.0) # :204 (first sub-frame)
    return x, am
