"""Host-side file plumbing around the encode path: WAV reading, segmentation rules, token writers.

Mirrors the behaviour (not the code) of the reference's
  * ``read_audio`` / ``convert_audio`` ............ audiotoken/utils.py:26-68  (mono mix-down, resample)
  * ``AudioBatchDataset._iter_chunk`` ............. audiotoken/datasets.py:75-105 (chunking, 3200-sample rule)
  * ``save_audio_tokens`` / ``save_rel_audio_tokens`` audiotoken/utils.py:199-225, 367-396 (on-disk layout)
  * ``sanitize_path`` / ``find_audio_files`` ...... audiotoken/utils.py:342-353, 172-184

On-disk contract: ``<outdir>/<name>.npy``, dtype int16, C order, shape ``(K, T_total)``; the tokens of
successive chunks of one file are concatenated along axis 1, each chunk trimmed to
``ceil(chunk_seconds * token_rate)`` tokens.  Unlike the reference (np.load + hstack + np.save per
chunk) a file is written exactly once, through a temporary name and an atomic rename, so re-running is
idempotent.  Only RIFF/WAV input is decoded here (no ffmpeg / torchcodec offline).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .configs import AUDIO_EXTS, MIN_SEGMENT_SAMPLES, AudioConfig


def sanitize_path(path) -> str:
    """expanduser, make absolute, resolve, mkdir -p (reference utils.py:342-353)."""
    p = Path(path).expanduser()
    if not p.is_absolute():
        p = p.absolute()
    p = p.resolve()
    if not p.exists():
        p.mkdir(parents=True, exist_ok=True)
    return str(p)


def find_audio_files(folder) -> List[str]:
    out = []
    for root, _dirs, files in os.walk(folder):
        for f in files:
            if f.lower().endswith(AUDIO_EXTS):
                out.append(os.path.join(root, f))
    return sorted(out)


def wav_info(path) -> Tuple[int, int, int]:
    """(sample_rate, num_frames, num_channels) from the RIFF header only."""
    import wave
    with wave.open(str(path), 'rb') as w:
        return w.getframerate(), w.getnframes(), w.getnchannels()


def convert_audio(audio: torch.Tensor, sample_rate: int, target_sample_rate: int) -> torch.Tensor:
    """[C, L] -> mono [1, L'] at the target rate (reference utils.py:26-44)."""
    c = audio.shape[0]
    if c == 2:
        audio = audio.mean(-2, keepdim=True)
    elif c != 1:
        raise RuntimeError("Only mono or stereo audio is supported")
    if sample_rate != target_sample_rate:
        import torchaudio
        audio = torchaudio.transforms.Resample(sample_rate, target_sample_rate)(audio)
    return audio


def read_audio(path, model_sample_rate: int, device=None) -> torch.Tensor:
    """WAV file -> fp32 [1, L] in [-1, 1] at `model_sample_rate` (reference utils.py:47-68).
    With a CUDA `device`, PCM decode + mono mix-down + resampling run on the GPU (audiotoken_b200/ingest.py) and the
    result stays there; otherwise the host path below (torchaudio, as the reference) is used."""
    if not str(path).lower().endswith('.wav'):
        raise NotImplementedError(f'{path}: only .wav can be decoded offline (no ffmpeg/torchcodec in this image)')
    from scipy.io import wavfile
    sr, data = wavfile.read(str(path))
    if device is not None and torch.device(device).type == 'cuda' and data.dtype in (np.int16, np.float32):
        from . import ingest
        raw = torch.from_numpy(np.ascontiguousarray(data).reshape(data.shape[0], -1)).t()     # [C, L] view of [L, C]
        return ingest.convert_audio(raw, int(sr), model_sample_rate, device)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 1:
        x = x[None, :]
    else:
        x = x.T
    audio = torch.from_numpy(np.ascontiguousarray(x))
    assert audio.dim() == 2, f"Audio needs to be 2D array, provided {audio.dim()}D for {path}"
    return convert_audio(audio, int(sr), model_sample_rate)


def _riff_pcm(path):
    """(sample_rate, channels, numpy dtype, byte offset, byte count) of a plain PCM / IEEE-float RIFF file, or None for
    anything this light parser does not handle (WAVE_FORMAT_EXTENSIBLE, 24-bit, RF64 ...: scipy reads those)."""
    import struct
    with open(path, 'rb') as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b'RIFF' or head[8:12] != b'WAVE':
            return None
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                return None
            cid, size = hdr[:4], struct.unpack('<I', hdr[4:])[0]
            if cid == b'fmt ':
                raw = f.read(size + (size & 1))
                if size < 16:
                    return None
                tag, ch, rate, _, _, bits = struct.unpack('<HHIIHH', raw[:16])
                dt = {(1, 16): np.int16, (1, 32): np.int32, (1, 8): np.uint8, (3, 32): np.float32}.get((tag, bits))
                if dt is None or ch < 1:
                    return None
                fmt = (rate, ch, dt)
            elif cid == b'data':
                if fmt is None:
                    return None
                return fmt[0], fmt[1], fmt[2], f.tell(), size
            else:
                f.seek(size + (size & 1), 1)


def read_wav_raw(path) -> Tuple[int, np.ndarray]:
    """Host-only half of the batch reader: RIFF/WAV file -> (sample_rate, PCM array [L] or [L, C]) untouched.
    Runs in the reader threads of the streaming file loop: no CUDA calls, and for plain PCM16 / PCM32 / u8 / float32
    files the payload is one `np.fromfile` (a single C call that releases the GIL, so 16 readers do not starve the
    thread that launches kernels); other encodings go through scipy."""
    if not str(path).lower().endswith('.wav'):
        raise NotImplementedError(f'{path}: only .wav can be decoded offline (no ffmpeg/torchcodec in this image)')
    info = _riff_pcm(str(path))
    if info is not None:
        sr, ch, dt, offset, nbytes = info
        item = np.dtype(dt).itemsize
        avail = max(0, os.path.getsize(str(path)) - offset)
        count = min(nbytes, avail) // (item * ch) * ch
        data = np.fromfile(str(path), dtype=dt, count=count, offset=offset)
        if ch > 1:
            data = data.reshape(-1, ch)
    else:
        from scipy.io import wavfile
        sr, data = wavfile.read(str(path))
    if data.ndim == 2 and data.shape[1] != 1:
        raise AssertionError(f'Audio needs to be mono, provided {data.shape[1]} channels for {path}')
    return int(sr), data


def upload_pcm(data: np.ndarray, device, dst: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """PCM16 / float32 payload of one file -> device tensor [1, n] (None for the encodings the device path does not take).
    The streaming file loop calls this for every file of a window on a copy-only stream BEFORE any ingest kernel of the
    window is queued: a pageable host->device copy blocks the calling thread until everything queued before it on its
    stream has run, and an ingest kernel queued between two copies runs only when the encoder's persistent kernels let
    go of an SM (0.3-1 ms) — per file, that serialised the host behind the device.  `dst`: a uint8 device slice of at
    least data.nbytes (16-byte aligned) to copy into instead of allocating."""
    if data.dtype not in (np.int16, np.float32):
        return None
    src = torch.from_numpy(np.ascontiguousarray(data).reshape(1, -1))
    if dst is None:
        return src.to(device, non_blocking=True)
    view = dst[:src.numel() * src.element_size()].view(src.dtype).view(1, -1)
    view.copy_(src, non_blocking=True)
    return view


def chunk_output_lengths(sr: int, n_samples: int, model_sample_rate: int, chunk_size: int) -> List[int]:
    """Samples each streamed chunk of a file has after resampling (what convert_chunks will return)."""
    g = math.gcd(int(sr), int(model_sample_rate))
    orig, new = int(sr) // g, int(model_sample_rate) // g
    step = int(chunk_size * sr)
    return [int(math.ceil(new * min(step, n_samples - a) / orig)) for a in range(0, n_samples, step)]


def convert_chunks(sr: int, data: np.ndarray, model_sample_rate: int, chunk_size: int, device=None,
                   raw: Optional[torch.Tensor] = None, out_alloc=None) -> List[torch.Tensor]:
    """Second half of the reference's batch reader (utils.py:82-101): the PCM stream is cut into chunks of `chunk_size`
    seconds AT ITS OWN sample rate and every chunk is resampled on its own — so a file whose rate differs from the
    model's has filter edges at every chunk boundary.  Returns fp32 [1, L_i] chunks; on a CUDA `device` PCM decode and
    resampling run on the GPU (ingest.py) and the chunks stay there."""
    on_gpu = device is not None and torch.device(device).type == 'cuda' and data.dtype in (np.int16, np.float32)
    if on_gpu:
        from . import ingest
        if raw is None:                                    # `raw`: the payload already uploaded by upload_pcm
            raw = upload_pcm(data, device)
    else:
        if data.dtype == np.int16:
            x = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            x = data.astype(np.float32) / 2147483648.0
        elif data.dtype == np.uint8:
            x = (data.astype(np.float32) - 128.0) / 128.0
        else:
            x = data.astype(np.float32)
        raw = torch.from_numpy(np.ascontiguousarray(x).reshape(1, -1))
    step = int(chunk_size * sr)
    out = []
    for a in range(0, raw.shape[1], step):
        piece = raw[:, a:a + step]
        if on_gpu:
            # out_alloc(n) -> contiguous float32 [1, n] device tensor (a slice of the window's buffer in the file loop)
            dst = out_alloc(ingest.resampler(sr, model_sample_rate, device).out_len(piece.shape[1])) if out_alloc else None
            out.append(ingest.convert_audio(piece, sr, model_sample_rate, device, out=dst))
        else:
            out.append(convert_audio(piece, sr, model_sample_rate))
    return out


def read_audio_chunks(path, model_sample_rate: int, chunk_size: int, device=None) -> List[torch.Tensor]:
    """The reference's batch reader (utils.py:71-101) = read_wav_raw + convert_chunks."""
    sr, data = read_wav_raw(path)
    return convert_chunks(sr, data, model_sample_rate, chunk_size, device)


def write_wav(path, wave: torch.Tensor, sample_rate: int) -> None:
    """fp32 [L] or [1, L] in [-1, 1] -> PCM16 WAV (synthetic corpora for tests / benchmarks)."""
    from scipy.io import wavfile
    x = wave.reshape(-1).clamp(-1, 1).numpy()
    wavfile.write(str(path), sample_rate, np.round(x * 32767.0).astype(np.int16))


@dataclass
class Segment:
    """One independently encoded unit: a <= chunk_size slice of a file."""
    file_name: str
    chunk_index: int
    wave: torch.Tensor            # fp32 [n], n >= MIN_SEGMENT_SAMPLES
    config: AudioConfig


def iter_segments(wave: torch.Tensor, file_name: str, sample_rate: int, token_rate: int,
                  chunk_size: int) -> Iterator[Segment]:
    """Chunking rules of the reference's dataset for one decoded file `wave` [1, L].

    The reference first streams `chunk_size`-second chunks (utils.py:82-101), then cuts each chunk into
    segments of chunk_size*sr samples (datasets.py:88-105) — i.e. one segment per chunk — skips a
    segment shorter than 3200 samples and right-pads the rest; length_seconds is per chunk.
    """
    seg_len = int(chunk_size * sample_rate)
    total = wave.shape[-1]
    for ci, start in enumerate(range(0, total, seg_len)):
        seg = wave[0, start:start + seg_len]
        n = int(seg.shape[0])
        if n < MIN_SEGMENT_SAMPLES:
            continue
        cfg = AudioConfig(file_name=file_name, start_idx=0, end_idx=min(seg_len, n),
                          length_seconds=n / sample_rate, length_samples=n, model_token_rate=token_rate)
        yield Segment(file_name, ci, seg, cfg)


def token_path_flat(file_name: str, root_dir: str) -> str:
    """reference save_audio_tokens: <root>/<basename up to the FIRST '.'>.npy (utils.py:202-203)."""
    stem = file_name.split('/')[-1].split('.')[0]
    return os.path.join(root_dir, f'{stem}.npy')


def token_path_rel(file_name: str, root_dir: str, rel_dir: str) -> str:
    """reference save_rel_audio_tokens: <root>/<relative dir>/<stem>.npy (utils.py:374-382)."""
    rel = os.path.dirname(os.path.relpath(file_name, start=rel_dir))
    stem = os.path.splitext(os.path.basename(file_name))[0]
    return os.path.join(root_dir, rel, f'{stem}.npy')


def save_tokens_atomic(path: str, chunks: Sequence[np.ndarray]) -> None:
    """Concatenate per-chunk (K, T_i) int16 arrays along axis 1 and write <path> exactly once."""
    arr = np.ascontiguousarray(np.hstack([np.asarray(c, dtype=np.int16) for c in chunks]))
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    tmp = f'{path}.tmp.{os.getpid()}'
    with open(tmp, 'wb') as f:
        np.save(f, arr)
    os.replace(tmp, path)


def save_audio_tokens(tokens, audio_pointer: AudioConfig, root_dir: str) -> None:
    """Drop-in for the reference helper of the same name (utils.py:199-225): trim to length_tokens and
    APPEND to an existing file along axis 1."""
    t = tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else np.asarray(tokens)
    t = t[:, :audio_pointer.length_tokens]
    path = token_path_flat(audio_pointer.file_name, root_dir)
    prev = [np.load(path)] if os.path.exists(path) else []
    save_tokens_atomic(path, prev + [t])


def save_rel_audio_tokens(tokens, audio_pointer: AudioConfig, root_dir: str, rel_dir: str) -> None:
    """Drop-in for reference utils.py:367-396."""
    t = tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else np.asarray(tokens)
    t = t[:, :audio_pointer.length_tokens]
    path = token_path_rel(audio_pointer.file_name, root_dir, rel_dir)
    prev = [np.load(path)] if os.path.exists(path) else []
    save_tokens_atomic(path, prev + [t])
