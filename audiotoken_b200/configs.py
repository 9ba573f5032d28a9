"""Tokenizer names and per-tokenizer constants of the encode path.

Mirrors the values (not the code) of the reference's ``audiotoken/configs.py``:
``Tokenizers`` (:20-23), ``AcousticEncoderConfig`` (:33-39), ``HubertEncoderConfig``
(:49-59), ``Wav2VecBertConfig`` (:112-135) and ``AudioConfig`` (:190-218).  Unlike the
reference nothing is downloaded at import time: weight locations are plain optional
paths, and ``weights=None`` means "seeded synthetic weights" (benchmark / parity mode).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from enum import Enum
from typing import Optional

AUDIO_EXTS = ('.mp3', '.flac', '.wav', '.ogg', '.opus')
TAR_EXTS = ('.tar', '.tar.gz', '.tgz', '.tar.bz2', '.tbz', '.tar.xz', '.txz')
ZIP_EXTS = ('.zip', '.ZIP')

# A segment shorter than this many samples is skipped (reference datasets.py:95-97;
# the constant is 3200 at both 16 kHz and 24 kHz).
MIN_SEGMENT_SAMPLES = 3200


class Tokenizers(str, Enum):
    """Same member names/values as the reference StrEnum (configs.py:20-23)."""
    acoustic = "acoustic"
    semantic_s = "semantic_s"
    semantic_m = "semantic_m"

    def __str__(self) -> str:  # StrEnum behaviour
        return self.value


@dataclass
class EncoderConfig:
    model_id: str
    model_sample_rate: int
    model_token_rate: int
    pad_token: Optional[int]


@dataclass
class AcousticEncoderConfig(EncoderConfig):
    """EnCodec 24 kHz: 75 frames/s, n_q = floor(bandwidth*1000 / (10*75))."""
    model_id: str = 'encodec'
    model_sample_rate: int = 24_000
    bandwidth: float = 12
    model_token_rate: int = 75
    pad_token: Optional[int] = 0
    weights: Optional[str] = None

    @property
    def num_codebooks(self) -> int:
        return int(math.floor(self.bandwidth * 1000 / (10 * self.model_token_rate)))


@dataclass
class Wav2VecBertConfig(EncoderConfig):
    """semantic_m: w2v-BERT 2.0 (21-layer cut), hidden state 19, VQ 2048 x 1024."""
    model_id: str = 'w2vbert2_l21'
    model_sample_rate: int = 16_000
    model_token_rate: int = 50
    output_layer: int = 19
    codebook_size: int = 2048
    hidden_size: int = 1024
    pad_token: Optional[int] = 0
    weights: Optional[str] = None          # dir/file with HF-named tensors
    quantizer_path: Optional[str] = None   # state_dict with `_codebook.embed` [1,K,D]


@dataclass
class SemanticSConfig(EncoderConfig):
    """semantic_s as BASELINE.json defines it: the same w2v-BERT 2.0 stack with a
    shallower cut and a k-means (cdist + argmin, reference encoder.py:100-101) codebook
    of 1000 centroids.  (The reference's own semantic_s is mHuBERT)"""
    model_id: str = 'w2vbert2_l21'
    model_sample_rate: int = 16_000
    model_token_rate: int = 50
    output_layer: int = 11
    codebook_size: int = 1000
    hidden_size: int = 1024
    pad_token: Optional[int] = 0
    weights: Optional[str] = None
    quantizer_path: Optional[str] = None


@dataclass
class HubertEncoderConfig(EncoderConfig):
    """The reference's own semantic_s (configs.py:49-59): mHuBERT-base, hidden state 11, k-means 1000 centres."""
    model_id: str = 'voidful/mhubert-base'
    model_sample_rate: int = 16_000
    model_token_rate: int = 50
    output_layer: int = 11
    codebook_size: int = 1000
    hidden_size: int = 768
    pad_token: Optional[int] = 0
    weights: Optional[str] = None          # dir/file with HF HubertModel tensors
    quantizer_path: Optional[str] = None   # joblib scikit-learn KMeans (mhubert_base_vp_en_es_fr_it3_L11_km1000.bin)


@dataclass
class AudioConfig:
    """Per-segment metadata; `length_tokens` is the number of tokens that are saved
    (reference configs.py:213-218: ceil(length_seconds * model_token_rate))."""
    file_name: str
    start_idx: Optional[int] = None
    end_idx: Optional[int] = None
    length_seconds: Optional[float] = None
    length_samples: Optional[int] = None
    model_token_rate: Optional[int] = None

    @property
    def length_tokens(self) -> int:
        if self.model_token_rate is None or self.length_seconds is None:
            raise ValueError("Model token rate or length of the audio file is not provided")
        return math.ceil(self.length_seconds * self.model_token_rate)


def num_codebooks_to_bandwidth(num_codebooks: int) -> float:
    """reference utils.py:432-443"""
    return {2: 1.5, 4: 3, 8: 6, 16: 12}[num_codebooks]


def bandwidth_to_num_codebooks(bandwidth: float) -> int:
    """reference utils.py:418-429"""
    return {1.5: 2, 3: 4, 6: 8, 12: 16, 24: 32}[bandwidth]
