"""Oracle: the reference's own `semantic_s` = mHuBERT-base + k-means.  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch-CPU tensor maths on an HF-named state dict, what `HubertEncoder.__call__`
(reference audiotoken/encoder.py:88-108) runs:
  * waveform normalisation ..... `hubert_processor` (encoder.py:20-26) = HF `Wav2Vec2FeatureExtractor` with
    do_normalize: (x - mean) / sqrt(var + 1e-7) over the WHOLE file, before chunking (datasets.py:78-79)
  * feature encoder ............ transformers hubert/modeling_hubert.py:154-213: 7 bias-free Conv1d
    (k 10,3,3,3,3,2,2 / s 5,2,2,2,2,2,2, 512 ch), GroupNorm(512 groups) + GELU on the first, GELU on the rest.
    The GroupNorm statistics run over EVERY frame of the zero-padded chunk (datasets.py:99-103 pads each chunk to
    chunk_size), so — unlike the w2v-BERT path — the result depends on the padded length.
  * feature projection ......... :216-230 (LayerNorm(512) -> Linear 512 -> 768)
  * frame mask ................. :675-701 (`_get_feature_vector_attention_mask`), padded frames zeroed :430-433
  * positional conv ............ :45-93 (weight-normed Conv1d k=128, pad 64, groups 16, last frame dropped, GELU)
  * encoder .................... :408-470 (x + pos -> LayerNorm -> 12 post-LN layers; hidden_states[i] = input of
    layer i), layer :372-400, attention = plain scaled-dot-product with the additive padding mask :236-259
  * tail ....................... encoder.py:96-103: hidden_states[11] -> affine-free LayerNorm(768) ->
    `torch.cdist` against the k-means centres -> argmin -> int16

Parity: pinned against HF `HubertModel(HubertConfig())` + `Wav2Vec2FeatureExtractor()` run in the build container
(tests/golden/make_golden_hubert.py -> tests/golden/hubert.npz).  No CUDA kernels exist for this path yet; the
oracle and its fixtures are the first step of the row (tier rule: oracle before kernels).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
HEADS, EPS = 12, 1e-5


def processor_normalize(wave: torch.Tensor) -> torch.Tensor:
    """Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm on one un-padded array (reference encoder.py:20-26)."""
    x = wave.to(torch.float32)
    return (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)


def feat_lengths(n_samples: torch.Tensor) -> torch.Tensor:
    """modeling_hubert.py:675-688: frames after the 7 valid (un-padded) convolutions."""
    n = n_samples.clone()
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        n = torch.div(n - k, s, rounding_mode='floor') + 1
    return n


def feature_encoder(wave: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """wave [B, L] (zero-padded chunks) -> [B, T, 512]."""
    h = wave[:, None]
    for i, s in enumerate(CONV_STRIDE):
        h = F.conv1d(h, sd[f'feature_extractor.conv_layers.{i}.conv.weight'], stride=s)
        if i == 0:
            C = h.shape[1]
            h = F.group_norm(h, C, sd['feature_extractor.conv_layers.0.layer_norm.weight'],
                             sd['feature_extractor.conv_layers.0.layer_norm.bias'], eps=1e-5)
        h = F.gelu(h)
    return h.transpose(1, 2)


def frame_mask(n_frames: int, sample_mask: torch.Tensor) -> torch.Tensor:
    """[B, L] 0/1 sample mask -> [B, T] bool frame mask (modeling_hubert.py:690-701)."""
    out_len = feat_lengths(sample_mask.sum(-1).to(torch.long))
    return torch.arange(n_frames)[None, :] < out_len[:, None]


def pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / |v| with the norm over (out, in/groups) per kernel tap."""
    g = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original0']
    v = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original1']
    return g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], EPS)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + '.weight'], sd[prefix + '.bias'])


def hidden_states(wave: torch.Tensor, sample_mask: torch.Tensor, sd: Dict[str, torch.Tensor],
                  n_layers: int = 12) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """-> (hidden_states[0..n_layers] each [B, T, 768], frame mask [B, T] bool), as HF returns them with
    output_hidden_states=True."""
    feats = feature_encoder(wave.to(torch.float32), sd)
    B, T, _ = feats.shape
    fm = frame_mask(T, sample_mask)
    h = _lin(_ln(feats, sd, 'feature_projection.layer_norm'), sd, 'feature_projection.projection')
    h = h * fm[..., None]                                              # padded frames -> 0  (:430-433)
    add_mask = torch.zeros(B, 1, 1, T)
    add_mask.masked_fill_(~fm[:, None, None, :], torch.finfo(torch.float32).min)
    pos = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd['encoder.pos_conv_embed.conv.bias'], padding=64, groups=16)
    pos = F.gelu(pos[:, :, :-1]).transpose(1, 2)                       # even kernel: drop the last frame (:95-103)
    h = _ln(h + pos, sd, 'encoder.layer_norm')
    out = []
    D = h.shape[-1]
    hd = D // HEADS
    for i in range(n_layers):
        out.append(h)
        p = f'encoder.layers.{i}.'
        q = _lin(h, sd, p + 'attention.q_proj').view(B, T, HEADS, hd).transpose(1, 2)
        k = _lin(h, sd, p + 'attention.k_proj').view(B, T, HEADS, hd).transpose(1, 2)
        v = _lin(h, sd, p + 'attention.v_proj').view(B, T, HEADS, hd).transpose(1, 2)
        w = torch.softmax(q @ k.transpose(2, 3) * hd ** -0.5 + add_mask, dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, T, D)
        h = _ln(h + _lin(a, sd, p + 'attention.out_proj'), sd, p + 'layer_norm')
        ff = _lin(F.gelu(_lin(h, sd, p + 'feed_forward.intermediate_dense')), sd, p + 'feed_forward.output_dense')
        h = _ln(h + ff, sd, p + 'final_layer_norm')
    out.append(h)
    return out, fm


def tokens(hs: List[torch.Tensor], centres: torch.Tensor, output_layer: int = 11) -> torch.Tensor:
    """reference encoder.py:96-103: LayerNorm (no affine) -> cdist -> argmin -> int16 [B, 1, T]."""
    e = F.layer_norm(hs[output_layer], (hs[output_layer].shape[-1],), None, None, EPS)
    d = torch.cdist(e, centres.to(e.dtype))
    return torch.argmin(d, dim=-1, keepdim=True).transpose(1, 2).to(torch.int16)


def hidden_states_ragged(clips: List[torch.Tensor], padded_len: int, sd: Dict[str, torch.Tensor],
                         n_layers: int = 12) -> List[List[torch.Tensor]]:
    """The same result WITHOUT materialising the zero padding — the layout a packed GPU batch would use.

    clips: un-padded (already normalised) waveforms; padded_len: the chunk length the reference pads every clip to.
    Per clip: conv0 on the valid samples zero-extended by k0 - 1 (every frame that touches a real sample); the
    GroupNorm statistics divide by the frame count of the PADDED chunk (the padded frames are exact zeros: conv0 has
    no bias); the remaining valid convolutions only need the first feat_lengths(n) frames, whose receptive fields lie
    inside the clip; padded frames are zero after the projection and masked in attention, so the encoder runs on the
    valid frames alone, with the positional conv seeing zeros beyond them.  Returns hidden_states per clip, each
    [T_valid, 768]."""
    k0, s0 = CONV_KERNEL[0], CONV_STRIDE[0]
    n_pad_frames = (padded_len - k0) // s0 + 1
    out = []
    for w in clips:
        n = w.shape[-1]
        t_valid = int(feat_lengths(torch.tensor(n)))
        x = F.pad(w.to(torch.float32), (0, k0 - 1))[None, None]
        h = F.conv1d(x, sd['feature_extractor.conv_layers.0.conv.weight'], stride=s0)       # [1, 512, frames touching the clip]
        h = h[:, :, :n_pad_frames]
        mean = h.sum(-1, keepdim=True) / n_pad_frames
        var = (h * h).sum(-1, keepdim=True) / n_pad_frames - mean * mean
        h = (h - mean) * torch.rsqrt(var + 1e-5)
        h = h * sd['feature_extractor.conv_layers.0.layer_norm.weight'][None, :, None] + sd['feature_extractor.conv_layers.0.layer_norm.bias'][None, :, None]
        h = F.gelu(h)
        # frames of layer 0 whose receptive field lies inside the clip: exactly what the valid output frames read
        h = h[:, :, :(n - k0) // s0 + 1]
        for i in range(1, len(CONV_STRIDE)):
            h = F.gelu(F.conv1d(h, sd[f'feature_extractor.conv_layers.{i}.conv.weight'], stride=CONV_STRIDE[i]))
        feats = h.transpose(1, 2)[:, :t_valid]
        x = _lin(_ln(feats, sd, 'feature_projection.layer_norm'), sd, 'feature_projection.projection')
        pos = F.conv1d(x.transpose(1, 2), pos_conv_weight(sd), sd['encoder.pos_conv_embed.conv.bias'], padding=64, groups=16)
        pos = F.gelu(pos[:, :, :t_valid]).transpose(1, 2)
        x = _ln(x + pos, sd, 'encoder.layer_norm')
        hs = []
        D = x.shape[-1]
        hd = D // HEADS
        for i in range(n_layers):
            hs.append(x[0])
            p = f'encoder.layers.{i}.'
            q = _lin(x, sd, p + 'attention.q_proj').view(1, t_valid, HEADS, hd).transpose(1, 2)
            k = _lin(x, sd, p + 'attention.k_proj').view(1, t_valid, HEADS, hd).transpose(1, 2)
            v = _lin(x, sd, p + 'attention.v_proj').view(1, t_valid, HEADS, hd).transpose(1, 2)
            a = (torch.softmax(q @ k.transpose(2, 3) * hd ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(1, t_valid, D)
            x = _ln(x + _lin(a, sd, p + 'attention.out_proj'), sd, p + 'layer_norm')
            ff = _lin(F.gelu(_lin(x, sd, p + 'feed_forward.intermediate_dense')), sd, p + 'feed_forward.output_dense')
            x = _ln(x + ff, sd, p + 'final_layer_norm')
        hs.append(x[0])
        out.append(hs)
    return out
