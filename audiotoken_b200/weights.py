"""Seeded synthetic weights with the tensor names of the real checkpoints.

Names and shapes follow the HF ``Wav2Vec2BertModel`` state dict that the reference loads
(``audiotoken/encoder.py:129``; shapes listed in SURVEY.md A.3), the VQ ``state_dict``
layout the reference reads (``_codebook.embed`` ``[1, K, D]``, ``audiotoken/utils.py:331-339``)
and, for the acoustic path, the HF ``EncodecModel`` encoder/quantizer names
(SURVEY.md A.7).  A real checkpoint with the same names loads through the same code.

All tensors are drawn on the CPU from one ``torch.Generator`` in a fixed order, so the GPU
box and the build container produce bit-identical weights for a given seed.
"""
from __future__ import annotations

from typing import Dict

import torch

W2VBERT = dict(hidden=1024, heads=16, head_dim=64, ffn=4096, feat_in=160,
               conv_kernel=31, left=64, right=8, ln_eps=1e-5)


def _randn(g, *shape, std=1.0, mean=0.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean


def synthetic_w2vbert_state_dict(n_layers: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """HF-named fp32 tensors for `n_layers` conformer layers + the feature projection.

    Scales follow HF's default init (Linear N(0, 0.02), distance embedding N(0, 1),
    pointwise conv ~N(0, 0.044), depthwise ~N(0, 0.25)); LayerNorm gains/biases and Linear
    biases are perturbed away from 1/0 so that parity tests exercise them.
    """
    g = torch.Generator().manual_seed(seed)
    H, F, D = W2VBERT['hidden'], W2VBERT['ffn'], W2VBERT['feat_in']
    sd: Dict[str, torch.Tensor] = {}

    def ln(prefix, n):
        sd[prefix + '.weight'] = _randn(g, n, std=0.1, mean=1.0)
        sd[prefix + '.bias'] = _randn(g, n, std=0.1)

    def lin(prefix, n_out, n_in, std=0.02, bias=True):
        sd[prefix + '.weight'] = _randn(g, n_out, n_in, std=std)
        if bias:
            sd[prefix + '.bias'] = _randn(g, n_out, std=0.02)

    ln('feature_projection.layer_norm', D)
    lin('feature_projection.projection', H, D, std=0.045)
    for i in range(n_layers):
        p = f'encoder.layers.{i}.'
        ln(p + 'ffn1_layer_norm', H)
        lin(p + 'ffn1.intermediate_dense', F, H)
        lin(p + 'ffn1.output_dense', H, F)
        ln(p + 'self_attn_layer_norm', H)
        for nm in ('linear_q', 'linear_k', 'linear_v', 'linear_out'):
            lin(p + 'self_attn.' + nm, H, H)
        sd[p + 'self_attn.distance_embedding.weight'] = _randn(
            g, W2VBERT['left'] + W2VBERT['right'] + 1, W2VBERT['head_dim'], std=1.0)
        ln(p + 'conv_module.layer_norm', H)
        sd[p + 'conv_module.pointwise_conv1.weight'] = _randn(g, 2 * H, H, 1, std=0.044)
        sd[p + 'conv_module.depthwise_conv.weight'] = _randn(g, H, 1, W2VBERT['conv_kernel'], std=0.25)
        ln(p + 'conv_module.depthwise_layer_norm', H)
        sd[p + 'conv_module.pointwise_conv2.weight'] = _randn(g, H, H, 1, std=0.044)
        ln(p + 'ffn2_layer_norm', H)
        lin(p + 'ffn2.intermediate_dense', F, H)
        lin(p + 'ffn2.output_dense', H, F)
        ln(p + 'final_layer_norm', H)
    return sd


def synthetic_codebook(codebook_size: int, dim: int, seed: int = 4) -> torch.Tensor:
    """`[K, D]` fp32 centroids ~ N(0, 1) (BASELINE config 5 recipe)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(codebook_size, dim, generator=g, dtype=torch.float32)


def data_derived_codebook(embeddings: torch.Tensor, codebook_size: int, seed: int = 2,
                          noise: float = 0.3) -> torch.Tensor:
    """Centroids sampled from LayerNormed embeddings + noise (random N(0,1)
    centroids collapse to ~140 used codes; data-derived ones give realistic margins)."""
    g = torch.Generator().manual_seed(seed)
    e = embeddings.reshape(-1, embeddings.shape[-1]).float().cpu()
    idx = torch.randint(0, e.shape[0], (codebook_size,), generator=g)
    return e[idx] + noise * torch.randn(codebook_size, e.shape[1], generator=g)


def synthetic_waveform(index: int, num_samples: int, sample_rate: int) -> torch.Tensor:
    """Clip `index` of the synthetic corpus: noise + 3 sinusoids, in [-1, 1]."""
    g = torch.Generator().manual_seed(1000 + index)
    t = torch.arange(num_samples, dtype=torch.float64) / sample_rate
    x = 0.1 * torch.randn(num_samples, generator=g, dtype=torch.float32)
    f = torch.rand(3, generator=g) * (4000.0 - 80.0) + 80.0
    a = torch.rand(3, generator=g) * (0.2 - 0.05) + 0.05
    ph = torch.rand(3, generator=g) * 2 * torch.pi
    for k in range(3):
        x = x + (a[k].double() * torch.sin(2 * torch.pi * f[k].double() * t + ph[k].double())).float()
    return x.clamp_(-1.0, 1.0)


# ---------------------------------------------------------------------------------------------------
# EnCodec 24 kHz SEANet encoder + RVQ codebooks (HF `EncodecModel` state-dict names)
# ---------------------------------------------------------------------------------------------------
# (name, C_in, C_out, kernel, stride): the 15 weight-normed causal convs in forward order
SEANET_CONVS = [
    ('encoder.layers.0.conv', 1, 32, 7, 1),
    ('encoder.layers.1.block.1.conv', 32, 16, 3, 1), ('encoder.layers.1.block.3.conv', 16, 32, 1, 1),
    ('encoder.layers.1.shortcut.conv', 32, 32, 1, 1), ('encoder.layers.3.conv', 32, 64, 4, 2),
    ('encoder.layers.4.block.1.conv', 64, 32, 3, 1), ('encoder.layers.4.block.3.conv', 32, 64, 1, 1),
    ('encoder.layers.4.shortcut.conv', 64, 64, 1, 1), ('encoder.layers.6.conv', 64, 128, 8, 4),
    ('encoder.layers.7.block.1.conv', 128, 64, 3, 1), ('encoder.layers.7.block.3.conv', 64, 128, 1, 1),
    ('encoder.layers.7.shortcut.conv', 128, 128, 1, 1), ('encoder.layers.9.conv', 128, 256, 10, 5),
    ('encoder.layers.10.block.1.conv', 256, 128, 3, 1), ('encoder.layers.10.block.3.conv', 128, 256, 1, 1),
    ('encoder.layers.10.shortcut.conv', 256, 256, 1, 1), ('encoder.layers.12.conv', 256, 512, 16, 8),
    ('encoder.layers.15.conv', 512, 128, 7, 1),
]


# EnCodec 24 kHz decoder (HF EncodecDecoder layer indices): (name, C_in, C_out, k, stride, transposed)
SEANET_DEC_CONVS = [
    ('decoder.layers.0.conv', 128, 512, 7, 1, False),
    ('decoder.layers.3.conv', 512, 256, 16, 8, True),
    ('decoder.layers.4.block.1.conv', 256, 128, 3, 1, False), ('decoder.layers.4.block.3.conv', 128, 256, 1, 1, False),
    ('decoder.layers.4.shortcut.conv', 256, 256, 1, 1, False),
    ('decoder.layers.6.conv', 256, 128, 10, 5, True),
    ('decoder.layers.7.block.1.conv', 128, 64, 3, 1, False), ('decoder.layers.7.block.3.conv', 64, 128, 1, 1, False),
    ('decoder.layers.7.shortcut.conv', 128, 128, 1, 1, False),
    ('decoder.layers.9.conv', 128, 64, 8, 4, True),
    ('decoder.layers.10.block.1.conv', 64, 32, 3, 1, False), ('decoder.layers.10.block.3.conv', 32, 64, 1, 1, False),
    ('decoder.layers.10.shortcut.conv', 64, 64, 1, 1, False),
    ('decoder.layers.12.conv', 64, 32, 4, 2, True),
    ('decoder.layers.13.block.1.conv', 32, 16, 3, 1, False), ('decoder.layers.13.block.3.conv', 16, 32, 1, 1, False),
    ('decoder.layers.13.shortcut.conv', 32, 32, 1, 1, False),
    ('decoder.layers.15.conv', 32, 1, 7, 1, False),
]


def synthetic_encodec_state_dict(seed: int = 0, n_codebooks: int = 32) -> Dict[str, torch.Tensor]:
    """HF-named fp32 tensors of the EnCodec 24 kHz encoder, its 2-layer LSTM and the RVQ codebooks."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, cin, cout, k, _s in SEANET_CONVS:
        v = _randn(g, cout, cin, k, std=1.0 / (cin * k) ** 0.5)
        norm = v.flatten(1).norm(dim=1).view(cout, 1, 1)
        sd[name + '.parametrizations.weight.original0'] = norm * (1.0 + 0.1 * _randn(g, cout, 1, 1))
        sd[name + '.parametrizations.weight.original1'] = v
        sd[name + '.bias'] = _randn(g, cout, std=0.05)
    bound = 1.0 / 512 ** 0.5
    for layer in range(2):
        for nm, shape in (('weight_ih', (2048, 512)), ('weight_hh', (2048, 512)), ('bias_ih', (2048,)), ('bias_hh', (2048,))):
            sd[f'encoder.layers.13.lstm.{nm}_l{layer}'] = (torch.rand(*shape, generator=g) * 2 - 1) * bound
    gq = torch.Generator().manual_seed(seed + 1)
    for q in range(n_codebooks):
        # later stages quantise smaller residuals: shrink the codebooks geometrically so every stage stays informative
        sd[f'quantizer.layers.{q}.codebook.embed'] = torch.randn(1024, 128, generator=gq) * (0.12 * 0.85 ** q)
    # decoder (own generator: the encoder / codebook tensors above do not depend on it)
    gd = torch.Generator().manual_seed(seed + 2)
    for name, cin, cout, k, _s, transposed in SEANET_DEC_CONVS:
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        v = _randn(gd, *shape, std=1.0 / (cin * k) ** 0.5 * (2.0 ** 0.5 if transposed else 1.0))
        norm = v.flatten(1).norm(dim=1).view(shape[0], 1, 1)
        sd[name + '.parametrizations.weight.original0'] = norm * (1.0 + 0.1 * _randn(gd, shape[0], 1, 1))
        sd[name + '.parametrizations.weight.original1'] = v
        sd[name + '.bias'] = _randn(gd, cout, std=0.05)
    for layer in range(2):
        for nm, shape in (('weight_ih', (2048, 512)), ('weight_hh', (2048, 512)), ('bias_ih', (2048,)), ('bias_hh', (2048,))):
            sd[f'decoder.layers.1.lstm.{nm}_l{layer}'] = (torch.rand(*shape, generator=gd) * 2 - 1) * bound
    return sd


def weight_norm_weight(sd: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    """w = g * v / |v| with the norm over (C_in, k) per output channel (torch weight_norm, dim=0)."""
    g_, v = sd[name + '.parametrizations.weight.original0'], sd[name + '.parametrizations.weight.original1']
    return g_ * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


# ---- mHuBERT-base (the reference's own `semantic_s`) -------------------------------
HUBERT = dict(hidden=768, heads=12, head_dim=64, ffn=3072, layers=12, conv_dim=512,
              conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2),
              pos_kernel=128, pos_groups=16, ln_eps=1e-5)


def synthetic_hubert_state_dict(seed: int = 0, n_layers: int = 12) -> Dict[str, torch.Tensor]:
    """HF ``HubertModel`` names and shapes (``voidful/mhubert-base`` = the base config the reference loads at
    ``audiotoken/encoder.py:72``): 7 bias-free strided convs with GroupNorm on the first, feature projection,
    weight-normed grouped positional conv, ``n_layers`` post-LN transformer layers.  Conv weights use the He-style
    scale of HF's init, Linear N(0, 0.02); norm gains/biases and Linear biases are perturbed away from 1/0."""
    g = torch.Generator().manual_seed(seed)
    H, F, C = HUBERT['hidden'], HUBERT['ffn'], HUBERT['conv_dim']
    sd: Dict[str, torch.Tensor] = {}

    def ln(prefix, n):
        sd[prefix + '.weight'] = _randn(g, n, std=0.1, mean=1.0)
        sd[prefix + '.bias'] = _randn(g, n, std=0.1)

    def lin(prefix, n_out, n_in, std=0.02):
        sd[prefix + '.weight'] = _randn(g, n_out, n_in, std=std)
        sd[prefix + '.bias'] = _randn(g, n_out, std=0.02)

    cin = 1
    for i, k in enumerate(HUBERT['conv_kernel']):
        sd[f'feature_extractor.conv_layers.{i}.conv.weight'] = _randn(g, C, cin, k, std=(2.0 / (cin * k)) ** 0.5)
        cin = C
    ln('feature_extractor.conv_layers.0.layer_norm', C)
    ln('feature_projection.layer_norm', C)
    lin('feature_projection.projection', H, C, std=0.03)
    pk, pg = HUBERT['pos_kernel'], HUBERT['pos_groups']
    v = _randn(g, H, H // pg, pk, std=2.0 * (1.0 / (pk * H)) ** 0.5)
    sd['encoder.pos_conv_embed.conv.parametrizations.weight.original1'] = v
    # weight_norm(dim=2): one gain per kernel tap, norm over (out, in/groups)
    sd['encoder.pos_conv_embed.conv.parametrizations.weight.original0'] = (
        v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt() * (1.0 + 0.1 * _randn(g, 1, 1, pk)))
    sd['encoder.pos_conv_embed.conv.bias'] = _randn(g, H, std=0.02)
    ln('encoder.layer_norm', H)
    for i in range(n_layers):
        p = f'encoder.layers.{i}.'
        for nm in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
            lin(p + 'attention.' + nm, H, H)
        ln(p + 'layer_norm', H)
        lin(p + 'feed_forward.intermediate_dense', F, H)
        lin(p + 'feed_forward.output_dense', H, F)
        ln(p + 'final_layer_norm', H)
    return sd
