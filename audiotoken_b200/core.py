"""Public API: ``AudioToken(tokenizer, device).encode / encode_batch_files``.

Same constructor and method signatures as the reference (audiotoken/core.py:27-289); the decode half
(audiotoken/core.py:291-359) is out of scope of this build.  Differences that do not change results:
  * weights are resolved lazily (nothing is downloaded at import time); without a checkpoint path the
    encoders use seeded synthetic weights of the named architecture;
  * ``encode_batch_files`` packs the segments of all files into ragged length-bucketed batches instead
    of padding every segment to ``chunk_size`` seconds, reads files in a thread pool, copies tokens to
    the host once per batch and writes every ``.npy`` exactly once (atomic rename);
  * ``device`` must be an sm_100 CUDA device — there is no CPU path.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import io as aio
from .configs import (AcousticEncoderConfig, AUDIO_EXTS, EncoderConfig, SemanticSConfig, Tokenizers,
                      Wav2VecBertConfig, num_codebooks_to_bandwidth)
from .packing import bucket_by_rows, length_tokens, padded_rows


class AudioToken:
    def __init__(self, tokenizer: Union[Tokenizers, str], device: str = "cuda:0", compile: bool = False, **kwargs):
        self.tokenizer_name = Tokenizers(tokenizer)
        self.encoder: Optional[torch.nn.Module] = None
        self.decoder = None
        self.model_config: EncoderConfig
        self.transform_func = None
        self.compile = compile          # accepted for signature compatibility; kernels are ahead-of-time compiled
        self.kwargs = kwargs
        self.device = device
        self.num_codebooks = kwargs.get("num_codebooks", 16)
        assert self.num_codebooks in [2, 4, 8, 16], "num_codebooks must be one of [2, 4, 8, 16]"
        self.load_config()

    # reference core.py:73-90
    def load_config(self):
        if self.tokenizer_name == Tokenizers.acoustic:
            self.model_config = AcousticEncoderConfig(bandwidth=num_codebooks_to_bandwidth(self.num_codebooks))
        elif self.tokenizer_name == Tokenizers.semantic_s:
            self.model_config = SemanticSConfig()
        elif self.tokenizer_name == Tokenizers.semantic_m:
            self.model_config = Wav2VecBertConfig()
        else:
            raise ValueError(f"Tokenizer {self.tokenizer_name} not supported")
        self.model_sample_rate = self.model_config.model_sample_rate

    # reference core.py:92-118
    def load_encoder(self):
        if self.encoder is not None:
            return
        enc_kw = {k: v for k, v in self.kwargs.items() if k in ('state_dict', 'codebook', 'precision', 'n_layers', 'seed')}
        if self.tokenizer_name == Tokenizers.acoustic:
            from .acoustic import AcousticEncoder
            self.encoder = AcousticEncoder(config=self.model_config, device=self.device, **enc_kw)
        elif self.tokenizer_name == Tokenizers.semantic_s:
            from .encoder import SemanticSEncoder
            self.encoder = SemanticSEncoder(config=self.model_config, device=self.device, **enc_kw)
        else:
            from .encoder import Wav2VecBertEncoder
            self.encoder = Wav2VecBertEncoder(config=self.model_config, device=self.device, quantize=True, **enc_kw)

    # reference core.py:120-185
    def encode(self, audio, chunk_size: Optional[int] = None) -> torch.Tensor:
        """ndarray / Tensor [1, L] at model_sample_rate, or a path -> int16 CPU tokens [1, K, T]
        ([K, T_total] when a path is encoded with chunk_size, as in the reference)."""
        self.load_encoder()
        if isinstance(audio, np.ndarray):
            assert audio.ndim == 2, "Audio must be 2D array"
            assert audio.shape[0] == 1, "Audio must mono"
            return self._encode_single(torch.from_numpy(audio))
        if isinstance(audio, torch.Tensor):
            assert audio.ndim == 2, "Audio must be 2D array"
            assert audio.shape[0] == 1, "Audio must mono"
            return self._encode_single(audio)
        if isinstance(audio, (os.PathLike, Path, str)) and not isinstance(audio, bytes):
            wave = aio.read_audio(audio, self.model_sample_rate)
            if chunk_size is None:
                return self._encode_single(wave)
            seg = int(chunk_size * self.model_sample_rate)
            parts = [self._encode_single(wave[:, s:s + seg])[0] for s in range(0, wave.shape[-1], seg)]
            return torch.cat(parts, dim=-1)
        if isinstance(audio, bytes):
            raise NotImplementedError("Encoding bytes not supported yet")
        raise ValueError(f"Unsupported input type {type(audio)}. Should be one of: ndarray, Tensor, PathLike")

    # reference core.py:187-196
    def _encode_single(self, audio: torch.Tensor) -> torch.Tensor:
        input_batch = audio.to(self.device, torch.float32)
        attention_mask = torch.ones_like(input_batch)
        toks = self.encoder(input_batch, attention_mask)
        return toks.cpu()

    # reference core.py:198-289
    def encode_batch_files(self, batch_size: int, outdir, chunk_size: int = 30, num_workers: int = 12,
                           audio_files: Optional[Sequence] = None, audio_dir=None, **dataloader_kwargs) -> None:
        self.load_encoder()
        assert audio_files or audio_dir, "Either audio_files or audio_dir must be provided"
        assert not (audio_files and audio_dir), "Provide either audio_files or audio_dir, not both"
        outdir = aio.sanitize_path(outdir)
        files = [str(f) for f in audio_files] if audio_files else aio.find_audio_files(str(audio_dir))
        num_workers = max(1, min(num_workers, os.cpu_count() or 1, len(files) or 1))
        sr, rate = self.model_sample_rate, self.model_config.model_token_rate
        self.last_stats = encode_files(self.encoder, files, outdir, sr, rate, chunk_size, batch_size, num_workers,
                                       rel_dir=None if audio_files else str(audio_dir))

    def load_decoder(self, **kwargs):
        """reference core.py:291-313.  Only the acoustic decoder is built."""
        if getattr(self, 'decoder', None) is not None:
            return
        if self.tokenizer_name != Tokenizers.acoustic:
            raise NotImplementedError("semantic token -> audio decoding (GPT-2 + Bark, reference decoder.py:79-) is outside "
                                      "the scope of this build (SURVEY.md section 8f)")
        from .acoustic import AcousticDecoder
        kw = {k: v for k, v in {**self.kwargs, **kwargs}.items() if k in ('state_dict', 'seed')}
        self.decoder = AcousticDecoder(device=self.device, **kw)

    def decode(self, tokens, **kwargs) -> torch.Tensor:
        """tokens (1, num_codebooks, num_tokens) -> audio fp32 CPU (1, num_samples)  (reference core.py:317-357)."""
        self.load_decoder(**kwargs)
        if isinstance(tokens, np.ndarray):
            tokens = torch.from_numpy(tokens)
        elif isinstance(tokens, (os.PathLike, str)):
            tokens = torch.load(tokens, map_location='cpu')
        elif not isinstance(tokens, torch.Tensor):
            raise ValueError(f"Unsupported input type {type(tokens)}. Should be one of: np.ndarray, torch.Tensor, os.PathLike")
        if tokens.dim() == 2:
            tokens = tokens.unsqueeze(0)
        return self.decoder(tokens.to(dtype=torch.long)).cpu()


def encode_files(encoder, files: Sequence[str], outdir: str, sample_rate: int, token_rate: int, chunk_size: int,
                 batch_size: int, num_workers: int, rel_dir: Optional[str]) -> Dict[str, float]:
    """The batched file loop: read -> segment -> ragged batches -> encode -> one .npy per file."""
    t0 = time.time()
    pad = int(chunk_size * sample_rate)
    row_budget = min(max(1, batch_size) * encoder.rows_for(pad), getattr(encoder, 'max_rows_per_batch', 1 << 30))

    def load(path):
        try:
            if not path.lower().endswith(AUDIO_EXTS):
                raise NotImplementedError(f'unsupported extension: {path}')
            chunks = aio.read_audio_chunks(path, sample_rate, chunk_size, device=getattr(encoder, 'device', None))
            segs_ = []
            for ci, wave in enumerate(chunks):                 # one segment per streamed chunk (datasets.py:88-105)
                for s in aio.iter_segments(wave, path, sample_rate, token_rate, chunk_size):
                    s.chunk_index = ci
                    segs_.append(s)
            return segs_
        except Exception as e:  # noqa: BLE001  (reference logs and continues, datasets.py:136-137)
            return e

    with ThreadPoolExecutor(num_workers) as ex:
        loaded = list(ex.map(load, files))
    segs, errors = [], {}
    for path, res in zip(files, loaded):
        if isinstance(res, Exception):
            errors[path] = repr(res)
        else:
            segs.extend(res)
    rows = [encoder.rows_for_tokens(length_tokens(int(s.wave.numel()), sample_rate, token_rate), pad) for s in segs]
    per_file: Dict[str, Dict[int, np.ndarray]] = {}
    audio_s = 0.0
    for idx in bucket_by_rows(rows, row_budget):
        clips = [segs[i].wave for i in idx]
        toks = encoder.encode_packed(clips, pad, [rows[i] for i in idx])
        host = [t.cpu().numpy() for t in toks]            # one sync per batch
        for i, t in zip(idx, host):
            s = segs[i]
            per_file.setdefault(s.file_name, {})[s.chunk_index] = t[:, :s.config.length_tokens]
            audio_s += s.config.length_seconds
    for path, chunks in per_file.items():
        dst = aio.token_path_flat(path, outdir) if rel_dir is None else aio.token_path_rel(path, outdir, rel_dir)
        aio.save_tokens_atomic(dst, [chunks[k] for k in sorted(chunks)])
    wall = time.time() - t0
    return {'files': len(per_file), 'segments': len(segs), 'audio_seconds': audio_s, 'wall_seconds': wall,
            'audio_seconds_per_second': audio_s / wall if wall > 0 else 0.0, 'errors': errors}
