"""Oracle: Wav2Vec2-BERT conformer stack with relative-key attention.  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch-CPU tensor maths on an HF-named state dict:
  * feature projection ........ transformers wav2vec2_bert/modeling_wav2vec2_bert.py:118-130
  * encoder entry / masks ..... same file :479-545 (padded rows zeroed :493, additive mask :496-500)
  * encoder layer ............. same file :422-460 (half-step FFN, MHSA, conv module, FFN, LN)
  * feed forward .............. same file :133-153
  * convolution module ........ same file :156-225
  * self attention ............ reference audiotoken/modeling_wav2vec2_bert.py:20-80
    (relative-key bias folded into an additive mask, SDPA with scale 1/8)
  * tail ...................... reference audiotoken/encoder.py:174-181 (hidden_states[L] ->
    affine-free LayerNorm -> quantiser)

``emulate_bf16=True`` reproduces the cast points of ``torch.amp.autocast('cuda', bfloat16)``
(reference encoder.py:164): Linear/conv/einsum/SDPA inputs and outputs are
rounded to bf16 with fp32 accumulation, LayerNorm/softmax run in fp32, the residual stream is
bf16 inside layer 0 and fp32 from layer 1 on.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

HEADS, HEAD_DIM, LEFT, RIGHT, KCONV, EPS = 16, 64, 64, 8, 31, 1e-5


def _rb(x: torch.Tensor, emu: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if emu else x


def _linear(x, w, b, emu):
    y = _rb(x, emu) @ _rb(w, emu).t()
    if b is not None:
        y = y + _rb(b, emu)
    return _rb(y, emu)


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, EPS)


def _swish(x, emu_round):
    return _rb(x * torch.sigmoid(x), emu_round)


def rel_key_attention(x, mask, sd, p, emu):
    """x [B, T, 1024] (LayerNormed), mask [B, T] 0/1 -> [B, T, 1024].

    reference audiotoken/modeling_wav2vec2_bert.py:37-77.
    """
    B, T, H = x.shape
    q = _linear(x, sd[p + 'linear_q.weight'], sd[p + 'linear_q.bias'], emu)
    k = _linear(x, sd[p + 'linear_k.weight'], sd[p + 'linear_k.bias'], emu)
    v = _linear(x, sd[p + 'linear_v.weight'], sd[p + 'linear_v.bias'], emu)
    q = q.view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
    k = k.view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
    v = v.view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
    pos = torch.arange(T)
    dist = (pos.view(1, -1) - pos.view(-1, 1)).clamp(-LEFT, RIGHT) + LEFT     # [T, T] (:49-52)
    E = _rb(sd[p + 'distance_embedding.weight'], emu)                         # [73, 64]
    r = _rb(torch.einsum('bhld,rd->bhlr', q, E), emu)                         # q . E_r
    bias = torch.gather(r, 3, dist.view(1, 1, T, T).expand(B, HEADS, T, T)) / 8.0   # (:57-58)
    pad = (1.0 - mask)[:, None, None, :] * torch.finfo(torch.bfloat16 if emu else torch.float32).min
    scores = torch.einsum('bhld,bhrd->bhlr', q, k) / 8.0 + (bias + pad)       # (:60-73)
    if emu:
        scores = scores.clamp(min=torch.finfo(torch.bfloat16).min)
    probs = torch.softmax(scores, dim=-1)
    o = _rb(torch.einsum('bhlr,bhrd->bhld', _rb(probs, emu), v), emu)
    o = o.transpose(1, 2).reshape(B, T, H)
    return _linear(o, sd[p + 'linear_out.weight'], sd[p + 'linear_out.bias'], emu)


def conv_module(x, mask, sd, p, emu):
    """transformers modeling_wav2vec2_bert.py:196-225."""
    h = _ln(x, sd[p + 'layer_norm.weight'], sd[p + 'layer_norm.bias'])
    h = h * mask.unsqueeze(-1)                                                # :200-201
    w1 = sd[p + 'pointwise_conv1.weight'][:, :, 0]                            # [2048, 1024]
    h = _linear(h, w1, None, emu)
    a, g = h[..., :1024], h[..., 1024:]
    h = _rb(a * torch.sigmoid(g), emu)                                        # GLU over channels
    wd = _rb(sd[p + 'depthwise_conv.weight'], emu)                            # [1024, 1, 31]
    hp = F.pad(h.transpose(1, 2), (KCONV - 1, 0))                             # causal left pad 30
    h = _rb(F.conv1d(hp, wd, groups=1024), emu).transpose(1, 2)
    h = _ln(h, sd[p + 'depthwise_layer_norm.weight'], sd[p + 'depthwise_layer_norm.bias'])
    h = h * torch.sigmoid(h)                                                  # swish, fp32
    w2 = sd[p + 'pointwise_conv2.weight'][:, :, 0]
    return _linear(h, w2, None, emu)


def ffn(x, sd, p, emu):
    h = _linear(x, sd[p + 'intermediate_dense.weight'], sd[p + 'intermediate_dense.bias'], emu)
    h = _swish(h, emu)
    return _linear(h, sd[p + 'output_dense.weight'], sd[p + 'output_dense.bias'], emu)


def encoder_layer(x, mask, sd, i, emu):
    """transformers modeling_wav2vec2_bert.py:422-460."""
    p = f'encoder.layers.{i}.'
    res_bf16 = emu and i == 0      # layer 0 receives the bf16 projection output
    h = _ln(x, sd[p + 'ffn1_layer_norm.weight'], sd[p + 'ffn1_layer_norm.bias'])
    x = _rb(ffn(h, sd, p + 'ffn1.', emu) * 0.5 + x, res_bf16)
    h = _ln(x, sd[p + 'self_attn_layer_norm.weight'], sd[p + 'self_attn_layer_norm.bias'])
    x = _rb(rel_key_attention(h, mask, sd, p + 'self_attn.', emu) + x, res_bf16)
    x = _rb(x + conv_module(x, mask, sd, p + 'conv_module.', emu), res_bf16)
    h = _ln(x, sd[p + 'ffn2_layer_norm.weight'], sd[p + 'ffn2_layer_norm.bias'])
    x = _rb(ffn(h, sd, p + 'ffn2.', emu) * 0.5 + x, res_bf16)
    return _ln(x, sd[p + 'final_layer_norm.weight'], sd[p + 'final_layer_norm.bias'])


def hidden_states(feats: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor],
                  n_layers: int, emulate_bf16: bool = False) -> List[torch.Tensor]:
    """input_features [B, T, 160], attention_mask [B, T] -> hidden_states[0..n_layers].

    hidden_states[i] = input of layer i = output of i layers (:511-513, :536-537).
    """
    emu = emulate_bf16
    feats = feats.float()
    mask = mask.float()
    h = _ln(feats, sd['feature_projection.layer_norm.weight'], sd['feature_projection.layer_norm.bias'])
    x = _linear(h, sd['feature_projection.projection.weight'], sd['feature_projection.projection.bias'], emu)
    x = x * mask.unsqueeze(-1)                                                # :493
    out = [x]
    for i in range(n_layers):
        x = encoder_layer(x, mask, sd, i, emu)
        out.append(x)
    return out


def final_embedding(h: torch.Tensor) -> torch.Tensor:
    """Affine-free LayerNorm before the quantiser (reference encoder.py:138-144, 175-176)."""
    return F.layer_norm(h.float(), (h.shape[-1],), None, None, EPS)
