"""Encoder modules behind ``AudioToken.encoder`` — host mirror of the reference's operator boundary.

The reference calls ``self.encoder(input_batch[B, L], attention_mask[B, L]) -> int16 [B, K, T]``
(audiotoken/core.py:194, :276); ``Wav2VecBertEncoder`` below keeps that call signature
(reference audiotoken/encoder.py:111-186) and additionally exposes ``encode_packed`` for ragged
batches.  All arithmetic happens in libb200tok.so; this file only prepares weights, plans the
batch and owns device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as L
from .configs import SemanticSConfig, Wav2VecBertConfig
from .fbank_tables import DeviceFbankTables
from .packing import DeviceBatch, SemanticPlan, plan_semantic
from .weights import synthetic_codebook, synthetic_w2vbert_state_dict

_PREC = {'bf16': L.PREC_BF16, 'fp32': L.PREC_FP32}


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class SemanticWeights:
    """HF-named state dict -> the preprocessed device tensors of pipeline.cu.

    bf16 precision mirrors ``torch.amp.autocast``: matmul weights are stored in bf16, biases are
    rounded to bf16 (autocast casts the bias too) but kept in fp32 storage, LayerNorm parameters
    stay fp32.  q/k/v projections are concatenated into one [3072, 1024] matrix and the GLU
    pointwise conv is row-interleaved so that both halves of a GLU pair land in the same tile.
    """

    def __init__(self, sd: Dict[str, torch.Tensor], n_layers: int, device, precision: str):
        self.tensors: Dict[str, torch.Tensor] = {}
        bf = precision == 'bf16'
        act = torch.bfloat16 if bf else torch.float32

        def W(x):
            return x.to(device=device, dtype=act).contiguous()

        def Bv(x):
            x = x.to(device=device, dtype=torch.float32)
            return (_bf16_round(x) if bf else x).contiguous()

        def Fv(x):
            return x.to(device=device, dtype=torch.float32).contiguous()

        t = self.tensors
        t['fp.ln.w'] = Fv(sd['feature_projection.layer_norm.weight'])
        t['fp.ln.b'] = Fv(sd['feature_projection.layer_norm.bias'])
        t['fp.proj.w'] = W(sd['feature_projection.projection.weight'])
        t['fp.proj.b'] = Bv(sd['feature_projection.projection.bias'])
        for i in range(n_layers):
            p, q = f'encoder.layers.{i}.', f'L{i}.'
            for f in ('ffn1', 'ffn2'):
                t[q + f + '.ln.w'] = Fv(sd[p + f + '_layer_norm.weight'])
                t[q + f + '.ln.b'] = Fv(sd[p + f + '_layer_norm.bias'])
                t[q + f + '.w1'] = W(sd[p + f + '.intermediate_dense.weight'])
                t[q + f + '.b1'] = Bv(sd[p + f + '.intermediate_dense.bias'])
                t[q + f + '.w2'] = W(sd[p + f + '.output_dense.weight'])
                t[q + f + '.b2'] = Bv(sd[p + f + '.output_dense.bias'])
            a = p + 'self_attn.'
            t[q + 'attn.ln.w'] = Fv(sd[p + 'self_attn_layer_norm.weight'])
            t[q + 'attn.ln.b'] = Fv(sd[p + 'self_attn_layer_norm.bias'])
            t[q + 'attn.wqkv'] = W(torch.cat([sd[a + 'linear_q.weight'], sd[a + 'linear_k.weight'],
                                              sd[a + 'linear_v.weight']], 0))
            t[q + 'attn.bqkv'] = Bv(torch.cat([sd[a + 'linear_q.bias'], sd[a + 'linear_k.bias'],
                                               sd[a + 'linear_v.bias']], 0))
            t[q + 'attn.wo'] = W(sd[a + 'linear_out.weight'])
            t[q + 'attn.bo'] = Bv(sd[a + 'linear_out.bias'])
            t[q + 'attn.dist'] = W(sd[a + 'distance_embedding.weight'])
            c = p + 'conv_module.'
            t[q + 'conv.ln.w'] = Fv(sd[c + 'layer_norm.weight'])
            t[q + 'conv.ln.b'] = Fv(sd[c + 'layer_norm.bias'])
            pw1 = sd[c + 'pointwise_conv1.weight'].reshape(2048, 1024)
            t[q + 'conv.pw1'] = W(torch.stack([pw1[:1024], pw1[1024:]], 1).reshape(2048, 1024))
            t[q + 'conv.dw'] = Fv(sd[c + 'depthwise_conv.weight'].reshape(1024, 31).t())   # tap-major [31, 1024]
            t[q + 'conv.dwln.w'] = Fv(sd[c + 'depthwise_layer_norm.weight'])
            t[q + 'conv.dwln.b'] = Fv(sd[c + 'depthwise_layer_norm.bias'])
            t[q + 'conv.pw2'] = W(sd[c + 'pointwise_conv2.weight'].reshape(1024, 1024))
            t[q + 'final.ln.w'] = Fv(sd[p + 'final_layer_norm.weight'])
            t[q + 'final.ln.b'] = Fv(sd[p + 'final_layer_norm.bias'])


class Wav2VecBertEncoder(torch.nn.Module):
    """semantic_m / semantic_s encoder: waveform -> int16 tokens [B, 1, T].

    Same constructor/call shape as reference audiotoken/encoder.py:111-186; extra keyword
    arguments select synthetic weights (``state_dict=None``) and the precision mode:
    ``'bf16'`` = the reference's CUDA autocast numerics on tcgen05 tensor cores,
    ``'fp32'`` = CUDA-core fp32 everywhere (the reference's CPU numerics; used for the 1e-4 check).
    """

    def __init__(self, config=None, device: str = 'cuda:0', quantize: bool = True,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 codebook: Optional[torch.Tensor] = None, precision: str = 'bf16',
                 n_layers: Optional[int] = None, seed: int = 0):
        super().__init__()
        self.config = config if config is not None else Wav2VecBertConfig()
        self.device = torch.device(device)
        L.require_device(self.device)
        self.lib = L.load()
        self.precision = precision
        self.quantize = quantize
        self.output_layer = self.config.output_layer
        self.n_layers = self.output_layer if n_layers is None else n_layers
        if state_dict is None:
            state_dict = synthetic_w2vbert_state_dict(self.n_layers, seed=seed)
        if codebook is None:
            codebook = synthetic_codebook(self.config.codebook_size, 1024, seed=4)
        self.codebook_size = int(codebook.shape[0])
        with torch.cuda.device(self.device):
            self.weights = SemanticWeights(state_dict, self.n_layers, self.device, precision)
            self.codebook = codebook.to(self.device, torch.float32).contiguous()
            self.tables = DeviceFbankTables(self.device)
            self.handle = self.lib.b2t_semantic_create(self.n_layers, self.codebook_size, _PREC[precision])
            if not self.handle:
                raise L.B2TError('b2t_semantic_create failed: ' + self.lib.b2t_last_error().decode())
            for name, t in self.weights.tensors.items():
                L.check(self.lib.b2t_semantic_set_tensor(self.handle, name.encode(), t.data_ptr()), name)
            L.check(self.lib.b2t_semantic_set_tensor(self.handle, b'codebook', self.codebook.data_ptr()), 'codebook')
        self._ws: Optional[torch.Tensor] = None
        self.last_launches = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.b2t_semantic_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_option(self, name: str, value: int) -> None:
        """'gemm_impl' / 'attn_impl' (lib.IMPL_*), 'mel_bf16' (0/1)."""
        L.check(self.lib.b2t_semantic_set_tensor(self.handle, f'opt.{name}'.encode(), C.c_void_p(value)), name)

    # ---- row bookkeeping used by the batched file loop -------------------------------------------
    num_codebooks = 1
    max_rows_per_batch = 262144

    def rows_for(self, padded_samples: int) -> int:
        """Rows the reference returns for a segment padded to `padded_samples` (T)."""
        from .packing import padded_rows
        return padded_rows(padded_samples)

    def rows_for_tokens(self, n_tokens: int, padded_samples: int) -> int:
        """Rows that have to be computed to save `n_tokens` tokens of a segment."""
        return max(1, min(n_tokens, self.rows_for(padded_samples)))

    # ---- packed (ragged) entry point ------------------------------------------------------------
    def _workspace(self, plan: SemanticPlan) -> torch.Tensor:
        need = self.lib.b2t_semantic_workspace_bytes(self.handle, plan.total_rows, plan.total_frames, plan.n_clips)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(int(need * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def encode_plan(self, wave: torch.Tensor, plan: SemanticPlan, tap_layer: int = -1
                    ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """wave: flat fp32 device tensor; returns tokens int16 [total_rows] (and the tapped
        hidden state fp32 [total_rows, 1024] when tap_layer >= 0)."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.is_contiguous()
        with torch.cuda.device(self.device):
            db = DeviceBatch(plan, self.device)
            ws = self._workspace(plan)
            tokens = torch.empty(plan.total_rows, dtype=torch.int16, device=self.device)
            tap = (torch.empty(plan.total_rows, 1024, dtype=torch.float32, device=self.device)
                   if tap_layer >= 0 else None)
            L.check(self.lib.b2t_semantic_encode(self.handle, wave.data_ptr(), db.byref(), self.tables.byref(),
                                                 ws.data_ptr(), ws.numel(), tokens.data_ptr(), tap_layer,
                                                 L.ptr(tap), L.stream_ptr()), 'b2t_semantic_encode')
            self.last_launches = self.lib.b2t_last_launch_count()
            self._keep = db   # keep the descriptor alive until the stream has consumed it
        return tokens, tap

    # ---- reference-shaped operator ---------------------------------------------------------------
    def forward(self, input_batch: torch.Tensor, mask: torch.Tensor, pad_to_multiple_of: int = 2,
                tap_layer: int = -1):
        """input_batch [B, L] fp32 on the device, mask [B, L] 0/1 right-padded -> int16 [B, 1, T]."""
        assert input_batch.dim() == 2, "Input tensor must have shape [batch, time]"
        B, Lp = input_batch.shape
        wave = input_batch.to(self.device, torch.float32).contiguous()
        lengths = mask.to(self.device).sum(dim=1).round().to(torch.int64).cpu().numpy()
        plan = plan_semantic(lengths, np.arange(B, dtype=np.int64) * Lp, Lp, None, pad_to_multiple_of)
        tokens, tap = self.encode_plan(wave.view(-1), plan, tap_layer)
        T = plan.total_rows // B
        out = tokens.view(B, 1, T)
        if tap_layer >= 0:
            return out, tap.view(B, T, 1024)
        return out

    def encode_packed(self, clips: Sequence[torch.Tensor], padded_samples, rows: Optional[Sequence[int]] = None
                      ) -> List[torch.Tensor]:
        """clips: list of 1-D fp32 tensors (any device) -> list of int16 [1, rows_i] device tensors."""
        lengths = [int(c.numel()) for c in clips]
        offs = np.zeros(len(clips), dtype=np.int64)
        offs[1:] = np.cumsum(lengths)[:-1]
        wave = torch.cat([c.reshape(-1).to(torch.float32) for c in clips]).to(self.device)
        plan = plan_semantic(lengths, offs, padded_samples, rows)
        tokens, _ = self.encode_plan(wave, plan)
        ro = plan.row_off
        return [tokens[ro[i]:ro[i + 1]].view(1, -1) for i in range(len(clips))]


class SemanticSEncoder(Wav2VecBertEncoder):
    """semantic_s as BASELINE.json defines it: shallower w2v-BERT cut + k-means assignment
    (`torch.cdist` + `argmin`, reference audiotoken/encoder.py:100-101) — same kernels."""

    def __init__(self, config=None, device: str = 'cuda:0', **kw):
        super().__init__(config if config is not None else SemanticSConfig(), device, True, **kw)
