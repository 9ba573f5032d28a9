"""Checkpoint loading for the encoders (reference audiotoken/encoder.py:38, :72-90, :132, :156-161).

The reference resolves its weights through ``hf_hub_download`` at import time (configs.py:55-133).  There is no
network here, so locations are plain paths: the ``weights`` / ``quantizer_path`` fields of the configs, or the
environment variables below.  Formats accepted are the ones the reference's files use:

  semantic_m   ``w2vbert2_l21/model.safetensors`` (HF ``Wav2Vec2BertModel`` names, optional ``wav2vec2_bert.`` prefix)
               + the VectorQuantize state dict ``..._ckpt8000.pkl`` (``torch.save``; ``_codebook.embed`` [1, K, D])
  semantic_s   ``mhubert_base_vp_en_es_fr_it3_L11_km1000.bin`` (joblib scikit-learn KMeans; ``cluster_centers_``)
  acoustic     the ``encodec`` package checkpoint (``encoder.model.N.conv.conv.weight_g/_v``,
               ``quantizer.vq.layers.N._codebook.embed``) or an HF ``EncodecModel`` state dict (current
               ``parametrizations.weight.original0/1`` or legacy ``weight_g/weight_v`` names)

Everything is returned with the HF names the weight preparation of encoder.py / acoustic.py consumes.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

import torch

ENV_W2VBERT = 'AUDIOTOKEN_W2VBERT_WEIGHTS'
ENV_VQ = 'AUDIOTOKEN_VQ_QUANTIZER'
ENV_ENCODEC = 'AUDIOTOKEN_ENCODEC_WEIGHTS'
ENV_KMEANS = 'AUDIOTOKEN_HUBERT_KMEANS'
ENV_HUBERT = 'AUDIOTOKEN_HUBERT_WEIGHTS'


def _load_tensor_file(path: str) -> Dict[str, torch.Tensor]:
    if os.path.isdir(path):
        for name in ('model.safetensors', 'pytorch_model.bin', 'model.pt', 'model.th'):
            p = os.path.join(path, name)
            if os.path.exists(p):
                path = p
                break
        else:
            raise FileNotFoundError(f'no model.safetensors / pytorch_model.bin under {path}')
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith('.safetensors'):
        from safetensors.torch import load_file
        return load_file(path, device='cpu')
    obj = torch.load(path, map_location='cpu', weights_only=True)
    if isinstance(obj, dict) and 'state_dict' in obj and isinstance(obj['state_dict'], dict):
        obj = obj['state_dict']
    if isinstance(obj, dict) and 'best_state' in obj and isinstance(obj['best_state'], dict):     # encodec training format
        obj = obj['best_state']
    if not isinstance(obj, dict):
        raise ValueError(f'{path}: expected a state dict, got {type(obj).__name__}')
    return {k: v for k, v in obj.items() if isinstance(v, torch.Tensor)}


def load_w2vbert_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """HF Wav2Vec2BertModel tensors (what ``Wav2Vec2BertModel.from_pretrained(config.model_id)`` reads, encoder.py:132)."""
    sd = _load_tensor_file(path)
    out = {}
    for k, v in sd.items():
        for prefix in ('wav2vec2_bert.', 'model.'):
            if k.startswith(prefix):
                k = k[len(prefix):]
        out[k] = v.float()
    if 'feature_projection.projection.weight' not in out:
        raise ValueError(f'{path}: not a Wav2Vec2BertModel state dict (feature_projection.projection.weight missing)')
    return out


def load_hubert_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """HF HubertModel tensors (``HubertModel.from_pretrained('voidful/mhubert-base')``, encoder.py:72)."""
    sd = _load_tensor_file(path)
    out = {(k[len('hubert.'):] if k.startswith('hubert.') else k): v.float() for k, v in sd.items()}
    if 'feature_extractor.conv_layers.0.conv.weight' not in out:
        raise ValueError(f'{path}: not a HubertModel state dict (feature_extractor.conv_layers.0.conv.weight missing)')
    return out


def w2vbert_num_layers(sd: Dict[str, torch.Tensor]) -> int:
    idx = [int(m.group(1)) for k in sd for m in [re.match(r'encoder\.layers\.(\d+)\.', k)] if m]
    return max(idx) + 1 if idx else 0


def load_vq_codebook(path: str) -> torch.Tensor:
    """VectorQuantize state dict (encoder.py:156-161, utils.py:331-338) -> codebook fp32 [K, D]."""
    sd = _load_tensor_file(path)
    for key in ('_codebook.embed', 'codebook.embed', 'embed', 'codebook'):
        if key in sd:
            cb = sd[key]
            break
    else:
        raise ValueError(f'{path}: no `_codebook.embed` tensor (keys: {sorted(sd)[:8]})')
    if cb.dim() == 3:
        if cb.shape[0] != 1:
            raise ValueError(f'{path}: multi-head codebook {tuple(cb.shape)} is not supported')
        cb = cb[0]
    return cb.float().contiguous()


def load_kmeans_centroids(path: str) -> torch.Tensor:
    """joblib scikit-learn KMeans (encoder.py:84-90) -> centroids fp32 [K, D]."""
    import joblib
    km = joblib.load(path)
    return torch.from_numpy(km.cluster_centers_).float().contiguous()


_ENCODEC_RULES = (
    (re.compile(r'^(encoder|decoder)\.model\.'), r'\1.layers.'),
    (re.compile(r'\.conv\.conv\.'), '.conv.'),
    (re.compile(r'\.convtr\.convtr\.'), '.conv.'),
    (re.compile(r'^quantizer\.vq\.layers\.(\d+)\._codebook\.'), r'quantizer.layers.\1.codebook.'),
    (re.compile(r'\.weight_g$'), '.parametrizations.weight.original0'),
    (re.compile(r'\.weight_v$'), '.parametrizations.weight.original1'),
)


def encodec_to_hf_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`encodec` package names / legacy HF weight_g, weight_v -> the HF EncodecModel names with parametrizations
    (the mapping of transformers' convert_encodec_checkpoint_to_pytorch: model.N -> layers.N, conv.conv -> conv,
    convtr.convtr -> conv, vq.layers.N._codebook -> layers.N.codebook)."""
    out = {}
    for k, v in sd.items():
        for rx, rep in _ENCODEC_RULES:
            k = rx.sub(rep, k)
        out[k] = v
    return out


def load_encodec_state_dict(path: str) -> Dict[str, torch.Tensor]:
    sd = encodec_to_hf_names(_load_tensor_file(path))
    if 'encoder.layers.0.conv.parametrizations.weight.original1' not in sd:
        raise ValueError(f'{path}: not an EnCodec 24 kHz state dict (encoder.layers.0.conv weight missing after renaming)')
    return {k: v.float() for k, v in sd.items()}


def resolve(path: Optional[str], env: str) -> Optional[str]:
    return path or os.environ.get(env) or None
