"""Oracle: EnCodec 24 kHz SEANet encoder + LSTM + residual VQ.  TEST INFRASTRUCTURE ONLY.

The reference calls the third-party `encodec` package (unpinned, requirements.txt:5; absent from
/root/reference): `self.model.encoder(x.unsqueeze(1))` then `self.model.quantizer.encode(emb, 75, bw)`
(reference audiotoken/encoder.py:44-57).  The algorithm is restated from the architecture that package
publishes and that transformers mirrors (models/encodec/modeling_encodec.py):
  * causal conv with reflect padding and the short-input rule ..... :82-176
  * residual block (ELU, k3, ELU, k1, + 1x1 shortcut) .............. :236-282
  * encoder layer order (ratios 2,4,5,8; LSTM x2 + skip; ELU; k7) .. :285-313
  * Euclidean codebook argmax(-(|r|^2 - 2 r.E^T + |E|^2)), residual update :364-438
tests/golden/make_golden.py runs HF `EncodecModel` (the stand-in SURVEY 8c names) with the same synthetic
weights and commits its embeddings and codes; tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from audiotoken_b200.weights import SEANET_CONVS, SEANET_DEC_CONVS, weight_norm_weight

_CONV = {name: (cin, cout, k, s) for name, cin, cout, k, s in SEANET_CONVS}
_CONV.update({name: (cin, cout, k, s) for name, cin, cout, k, s, _t in SEANET_DEC_CONVS})


def _pad1d_reflect(x: torch.Tensor, left: int, right: int) -> torch.Tensor:
    """modeling_encodec.py:139-155 (reflect pad with zero-extension of inputs shorter than the pad)."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    y = F.pad(x, (left, right), mode='reflect')
    return y[..., :y.shape[-1] - extra]


def conv(x: torch.Tensor, sd: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    """x [B, C_in, L] -> [B, C_out, ceil(L/s)]  (causal EncodecConv1d, weight-normed)."""
    _cin, _cout, k, s = _CONV[name]
    length = x.shape[-1]
    pad_total = k - s
    n_frames = math.ceil((length - k + pad_total) / s + 1) - 1
    extra = n_frames * s + k - pad_total - length
    x = _pad1d_reflect(x, pad_total, extra)
    return F.conv1d(x, weight_norm_weight(sd, name), sd[name + '.bias'], stride=s)


def lstm(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = 'encoder.layers.13.lstm.') -> torch.Tensor:
    """x [B, 512, T] -> LSTM(2 layers)(x) + x   (modeling_encodec.py:222-233; gate order i, f, g, o)."""
    h_in = x.permute(2, 0, 1)                                    # [T, B, 512]
    inp = h_in
    for layer in range(2):
        p = prefix
        w_ih, w_hh = sd[p + f'weight_ih_l{layer}'], sd[p + f'weight_hh_l{layer}']
        b = sd[p + f'bias_ih_l{layer}'] + sd[p + f'bias_hh_l{layer}']
        T, B, _ = inp.shape
        h = torch.zeros(B, 512, dtype=x.dtype)
        c = torch.zeros(B, 512, dtype=x.dtype)
        xg = inp @ w_ih.t() + b
        outs = []
        for t in range(T):
            gates = xg[t] + h @ w_hh.t()
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, 0)
    return (inp + h_in).permute(1, 2, 0)


def encoder(wave: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """wave [B, L] -> embeddings [B, 128, ceil(L/320)]  (reference encoder.py:48)."""
    x = conv(wave.unsqueeze(1).float(), sd, 'encoder.layers.0.conv')
    for blk, down in ((1, 3), (4, 6), (7, 9), (10, 12)):
        p = f'encoder.layers.{blk}.'
        h = conv(F.elu(x), sd, p + 'block.1.conv')
        h = conv(F.elu(h), sd, p + 'block.3.conv')
        x = conv(x, sd, p + 'shortcut.conv') + h
        x = conv(F.elu(x), sd, f'encoder.layers.{down}.conv')
    x = lstm(x, sd)
    return conv(F.elu(x), sd, 'encoder.layers.15.conv')


def rvq_codes(emb: torch.Tensor, sd: Dict[str, torch.Tensor], n_q: int) -> torch.Tensor:
    """emb [B, 128, T] -> codes int64 [n_q, B, T].  Residual arithmetic in fp32 (as the reference), the
    nearest codeword of each stage decided on exact (fp64) distances."""
    B, D, T = emb.shape
    r = emb.permute(0, 2, 1).reshape(-1, D).float().clone()
    out = []
    for q in range(n_q):
        E = sd[f'quantizer.layers.{q}.codebook.embed'].float()
        rd, Ed = r.double(), E.double()
        d = (rd * rd).sum(1, keepdim=True) - 2.0 * (rd @ Ed.t()) + (Ed * Ed).sum(1).unsqueeze(0)
        idx = torch.argmin(d, dim=1)
        out.append(idx.view(B, T))
        r = r - E[idx]
    return torch.stack(out, 0)


def rvq_codes_reference_fp32(emb: torch.Tensor, sd: Dict[str, torch.Tensor], n_q: int) -> torch.Tensor:
    """The reference's own fp32 expression verbatim (modeling_encodec.py:364-369, 424-438)."""
    B, D, T = emb.shape
    r = emb.permute(0, 2, 1).reshape(-1, D).float()
    out = []
    for q in range(n_q):
        E = sd[f'quantizer.layers.{q}.codebook.embed'].float()
        et = E.t()
        dist = -(r.pow(2).sum(1, keepdim=True) - 2 * r @ et + et.pow(2).sum(0, keepdim=True))
        idx = dist.max(dim=-1).indices
        out.append(idx.view(B, T))
        r = r - F.embedding(idx, E)
    return torch.stack(out, 0)


# ---- decode half (reference audiotoken/decoder.py:62-76: quantizer.decode + decoder) ---------------
def conv_transpose(x: torch.Tensor, sd: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    """x [B, C_in, T] -> [B, C_out, T*s]: causal EncodecConvTranspose1d (modeling_encodec.py:179-219): the k - s
    trailing samples of the full transposed convolution are trimmed (trim_right_ratio = 1)."""
    _cin, _cout, k, s = _CONV[name]
    y = F.conv_transpose1d(x, weight_norm_weight(sd, name), sd[name + '.bias'], stride=s)
    return y[..., :y.shape[-1] - (k - s)]


def rvq_decode(codes: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """codes int [n_q, B, T] -> embeddings [B, 128, T]: sum of the selected codewords in stage order
    (modeling_encodec.py:440-447 `quantized_out = quantized_out + quantized`)."""
    n_q, B, T = codes.shape
    out = torch.zeros(B, T, 128)
    for q in range(n_q):
        out = out + sd[f'quantizer.layers.{q}.codebook.embed'].float()[codes[q].long()]
    return out.permute(0, 2, 1)


def decoder(emb: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """emb [B, 128, T] -> waveform [B, 1, 320*T]  (modeling_encodec.py:316-353: conv k7, LSTM + skip, 4 x (ELU,
    transposed conv, residual block), ELU, conv k7)."""
    x = conv(emb.float(), sd, 'decoder.layers.0.conv')
    x = lstm(x, sd, 'decoder.layers.1.lstm.')
    for up, blk in ((3, 4), (6, 7), (9, 10), (12, 13)):
        x = conv_transpose(F.elu(x), sd, f'decoder.layers.{up}.conv')
        p = f'decoder.layers.{blk}.'
        h = conv(F.elu(x), sd, p + 'block.1.conv')
        h = conv(F.elu(h), sd, p + 'block.3.conv')
        x = conv(x, sd, p + 'shortcut.conv') + h
    return conv(F.elu(x), sd, 'decoder.layers.15.conv')
