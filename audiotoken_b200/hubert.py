"""The reference's own ``semantic_s``: mHuBERT-base + k-means (reference audiotoken/encoder.py:60-108, configs.py:49-59).

Host mirror of ``HubertEncoder``: same call shape (``encoder(input_batch[B, L], attention_mask[B, L]) -> int16
[B, 1, T]``) plus ``encode_packed`` for ragged batches; all arithmetic is in libb200tok.so (csrc/hubert.cu).  This file
plans the batch (per-level row tables, work lists), prepares the weights and owns device buffers.

Frames.  A clip of n samples has ``t_valid = feat_lengths(n)`` real frames; the reference pads every chunk to
``chunk_size`` and returns ``T = feat_lengths(L)`` frames, of which ``ceil(n / 320)`` are saved (configs.py:213-218) —
up to two more than ``t_valid``.  Those padded frames are well defined (zero after the projection, modeling_hubert.py
:430-433; masked as keys; still queries), so a clip is planned with ``rows >= t_valid`` token rows: the extra rows
enter the positional conv as zeros and the attention as queries only.  The GroupNorm of the first conv layer divides by
the frame count of the padded chunk (``padded_samples``), which is why this front end is not padding-invariant.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as L
from .configs import HubertEncoderConfig
from .packing import attention_work_lists

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
HID, FFN, HEADS, CONV_DIM, POS_K, POS_G = 768, 3072, 12, 512, 128, 16
TILE_F = 128


def feat_lengths(n):
    """frames after the 7 valid convolutions (modeling_hubert.py:675-688)"""
    n = np.asarray(n, dtype=np.int64).copy()
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        n = (n - k) // s + 1
    return n


class HubertBatchStruct(C.Structure):
    """b2t_hubert_batch"""
    _fields_ = [('n_clips', C.c_int32), ('total_rows', C.c_int32), ('level0_rows', C.c_int32), ('pos_rows', C.c_int32),
                ('n_stat_tiles', C.c_int32), ('n_apply_tiles', C.c_int32), ('total_samples', C.c_int64),
                ('wave_off', C.c_void_p), ('norm_off', C.c_void_p), ('n_samples', C.c_void_p), ('gn_count', C.c_void_p),
                ('off0', C.c_void_p), ('row_off', C.c_void_p), ('pos_off', C.c_void_p), ('stat_tile_clip', C.c_void_p),
                ('stat_tile_f0', C.c_void_p), ('stat_tile_first', C.c_void_p), ('apply_tile_clip', C.c_void_p),
                ('apply_tile_f0', C.c_void_p), ('attn', L.Batch)]


@dataclass
class HubertPlan:
    n_clips: int
    n_samples: np.ndarray       # int32 [n]
    wave_off: np.ndarray        # int64 [n]
    norm_off: np.ndarray        # int64 [n]
    gn_count: np.ndarray        # int32 [n]
    t_valid: np.ndarray         # int32 [n]
    rows: np.ndarray            # int32 [n]
    off0: np.ndarray            # int32 [n+1]
    row_off: np.ndarray         # int32 [n+1]
    pos_off: np.ndarray         # int32 [n]
    pos_rows: int
    stat_tile_clip: np.ndarray
    stat_tile_f0: np.ndarray
    stat_tile_first: np.ndarray
    apply_tile_clip: np.ndarray
    apply_tile_f0: np.ndarray
    qtile_clip: np.ndarray
    qtile_q0: np.ndarray
    qtile128_clip: np.ndarray
    qtile128_q0: np.ndarray

    @property
    def total_rows(self) -> int:
        return int(self.row_off[-1])

    @property
    def level0_rows(self) -> int:
        return int(self.off0[-1])


def plan_hubert(lengths: Sequence[int], wave_offsets: Sequence[int], padded_samples,
                rows: Optional[Sequence[int]] = None) -> HubertPlan:
    """lengths[i] valid samples of clip i (>= 400), wave_offsets[i] its first sample in the wave buffer,
    padded_samples the length the reference pads the chunk to (int or per clip), rows[i] token rows to compute
    (default: all T frames the reference returns)."""
    n = len(lengths)
    ln = np.asarray(lengths, dtype=np.int64)
    if n == 0 or np.any(ln < 400):
        raise ValueError('every clip needs at least 400 samples (one HuBERT frame)')
    pad = np.broadcast_to(np.asarray(padded_samples, dtype=np.int64), (n,))
    if np.any(pad < ln):
        raise ValueError('padded_samples must be >= the clip length')
    t_valid = feat_lengths(ln)
    t_pad = feat_lengths(pad)
    rows_arr = t_pad.copy() if rows is None else np.asarray(rows, dtype=np.int64)
    if np.any(rows_arr < t_valid) or np.any(rows_arr > t_pad):
        raise ValueError('rows[i] must be in [valid frames, frames of the padded chunk]')
    # feature-encoder levels: level 6 allots t_valid + 2 rows, level l twice level l + 1 (offsets halve exactly)
    a6 = t_valid + 2
    off6 = np.zeros(n + 1, dtype=np.int64)
    off6[1:] = np.cumsum(a6)
    off0 = off6 * 64
    n_l = (ln - CONV_KERNEL[0]) // CONV_STRIDE[0] + 1            # frames of conv0 that lie inside the clip
    n0_in = n_l.copy()
    for lvl in range(1, 7):
        assert np.all(n_l <= (a6 << (6 - (lvl - 1)))), 'level allotment too small'
        n_l = (n_l - CONV_KERNEL[lvl]) // 2 + 1
    assert np.array_equal(n_l, t_valid)
    gn_count = (pad - CONV_KERNEL[0]) // CONV_STRIDE[0] + 1
    n_stat = np.minimum((ln + CONV_STRIDE[0] - 1) // CONV_STRIDE[0], gn_count)
    row_off = np.zeros(n + 1, dtype=np.int64)
    row_off[1:] = np.cumsum(rows_arr)
    gap = POS_K // 2
    pos_off = np.zeros(n, dtype=np.int64)
    acc = gap
    for i in range(n):
        pos_off[i] = acc
        acc += int(rows_arr[i]) + gap
    pos_rows = acc + gap
    if off0[-1] >= 2 ** 31 or row_off[-1] >= 2 ** 31 or pos_rows * 16 * 128 >= 2 ** 31:
        raise ValueError('batch too large for int32 offsets')

    def tiles(counts):
        tc, tf, first = [], [], np.zeros(n + 1, dtype=np.int64)
        for i in range(n):
            f0 = np.arange(0, counts[i], TILE_F, dtype=np.int32)
            tc.append(np.full(f0.shape, i, dtype=np.int32))
            tf.append(f0)
            first[i + 1] = first[i] + f0.size
        return np.concatenate(tc), np.concatenate(tf), first.astype(np.int32)

    stc, stf, sfirst = tiles(n_stat)
    atc, atf, _ = tiles(n0_in)
    norm_off = np.zeros(n, dtype=np.int64)
    norm_off[1:] = np.cumsum(ln)[:-1]
    qc, qq, qc8, qq8 = attention_work_lists(rows_arr, t_valid)
    return HubertPlan(n_clips=n, n_samples=ln.astype(np.int32), wave_off=np.asarray(wave_offsets, dtype=np.int64),
                      norm_off=norm_off, gn_count=gn_count.astype(np.int32), t_valid=t_valid.astype(np.int32),
                      rows=rows_arr.astype(np.int32), off0=off0.astype(np.int32), row_off=row_off.astype(np.int32),
                      pos_off=pos_off.astype(np.int32), pos_rows=int(pos_rows), stat_tile_clip=stc, stat_tile_f0=stf,
                      stat_tile_first=sfirst, apply_tile_clip=atc, apply_tile_f0=atf, qtile_clip=qc, qtile_q0=qq,
                      qtile128_clip=qc8, qtile128_q0=qq8)


class DeviceHubertBatch:
    """A HubertPlan uploaded to the device (one pinned staging buffer, one copy) + the ctypes b2t_hubert_batch."""

    def __init__(self, plan: HubertPlan, device):
        self.plan = plan
        i64 = [plan.wave_off, plan.norm_off]
        i32 = [plan.n_samples, plan.gn_count, plan.off0, plan.row_off, plan.pos_off, plan.stat_tile_clip, plan.stat_tile_f0,
               plan.stat_tile_first, plan.apply_tile_clip, plan.apply_tile_f0, plan.t_valid, plan.qtile_clip, plan.qtile_q0,
               plan.qtile128_clip, plan.qtile128_q0]
        total = sum(2 * a.size for a in i64) + sum(a.size + (a.size & 1) for a in i32)
        host = torch.empty(total, dtype=torch.int32, pin_memory=torch.cuda.is_available())
        hv = host.numpy()
        off, offs = 0, []
        for a in i64:
            hv[off:off + 2 * a.size] = np.ascontiguousarray(a, dtype=np.int64).view(np.int32)
            offs.append(off)
            off += 2 * a.size
        for a in i32:
            hv[off:off + a.size] = a
            offs.append(off)
            off += a.size + (a.size & 1)
        self.host = host
        self.dev = host.to(device, non_blocking=True)
        base = self.dev.data_ptr()
        p = [base + 4 * o for o in offs]
        b = HubertBatchStruct()
        b.n_clips, b.total_rows, b.level0_rows, b.pos_rows = plan.n_clips, plan.total_rows, plan.level0_rows, plan.pos_rows
        b.n_stat_tiles, b.n_apply_tiles = int(plan.stat_tile_clip.size), int(plan.apply_tile_clip.size)
        b.total_samples = int(plan.n_samples.astype(np.int64).sum())
        (b.wave_off, b.norm_off, b.n_samples, b.gn_count, b.off0, b.row_off, b.pos_off, b.stat_tile_clip, b.stat_tile_f0,
         b.stat_tile_first, b.apply_tile_clip, b.apply_tile_f0, valid_rows, qc, qq, qc8, qq8) = p
        a = b.attn
        a.n_clips, a.total_frames, a.total_rows = plan.n_clips, 0, plan.total_rows
        a.n_qtiles, a.n_ctiles, a.max_rows = int(plan.qtile_clip.size), 0, int(plan.rows.max())
        a.row_off, a.valid_rows, a.qtile_clip, a.qtile_q0 = b.row_off, valid_rows, qc, qq
        a.n_qtiles128, a.qtile128_clip, a.qtile128_q0 = int(plan.qtile128_clip.size), qc8, qq8
        self.c = b

    def byref(self):
        return C.byref(self.c)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class HubertWeights:
    """HF ``HubertModel`` state dict -> the device tensors of csrc/hubert.cu.  bf16 mirrors torch.amp.autocast: matmul /
    conv weights in bf16, their biases rounded to bf16 (kept in fp32 storage), norm parameters fp32."""

    def __init__(self, sd: Dict[str, torch.Tensor], n_layers: int, device, precision: str):
        bf = precision == 'bf16'
        act = torch.bfloat16 if bf else torch.float32
        t: Dict[str, torch.Tensor] = {}

        def W(x):
            return x.to(device=device, dtype=act).contiguous()

        def Bv(x):
            x = x.to(device=device, dtype=torch.float32)
            return (_bf16_round(x) if bf else x).contiguous()

        def Fv(x):
            return x.to(device=device, dtype=torch.float32).contiguous()

        sd = {k[len('hubert.'):] if k.startswith('hubert.') else k: v for k, v in sd.items()}
        t['fe.conv0.w'] = Fv(sd['feature_extractor.conv_layers.0.conv.weight'].reshape(CONV_DIM, CONV_KERNEL[0]))
        t['fe.gn.w'] = Fv(sd['feature_extractor.conv_layers.0.layer_norm.weight'])
        t['fe.gn.b'] = Fv(sd['feature_extractor.conv_layers.0.layer_norm.bias'])
        for lvl in range(1, 7):
            w = sd[f'feature_extractor.conv_layers.{lvl}.conv.weight']               # [512, 512, k]
            t[f'fe.conv{lvl}.w'] = W(w.permute(0, 2, 1).reshape(CONV_DIM, -1))         # K index = tap * 512 + c_in
        t['fp.ln.w'] = Fv(sd['feature_projection.layer_norm.weight'])
        t['fp.ln.b'] = Fv(sd['feature_projection.layer_norm.bias'])
        t['fp.proj.w'] = W(sd['feature_projection.projection.weight'])
        t['fp.proj.b'] = Bv(sd['feature_projection.projection.bias'])
        # positional conv: weight_norm(dim=2) folded, group g -> [128 (48 real outputs), tap * 48 + c_in]
        if 'encoder.pos_conv_embed.conv.parametrizations.weight.original0' in sd:
            g_ = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original0'].float()
            v_ = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original1'].float()
        elif 'encoder.pos_conv_embed.conv.weight_g' in sd:
            g_, v_ = sd['encoder.pos_conv_embed.conv.weight_g'].float(), sd['encoder.pos_conv_embed.conv.weight_v'].float()
        else:
            g_, v_ = None, sd['encoder.pos_conv_embed.conv.weight'].float()
        w = v_ if g_ is None else g_ * v_ / v_.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()   # [768, 48, 128]
        gc = HID // POS_G
        pw = torch.zeros(POS_G, 128, POS_K * gc)
        pb = torch.zeros(POS_G, 128)
        for g in range(POS_G):
            pw[g, :gc] = w[g * gc:(g + 1) * gc].permute(0, 2, 1).reshape(gc, POS_K * gc)
            pb[g, :gc] = sd['encoder.pos_conv_embed.conv.bias'][g * gc:(g + 1) * gc]
        t['pos.w'] = W(pw)
        t['pos.b'] = Bv(pb)
        t['enc.ln.w'] = Fv(sd['encoder.layer_norm.weight'])
        t['enc.ln.b'] = Fv(sd['encoder.layer_norm.bias'])
        t['attn.zero_bias'] = torch.zeros(73, 64, dtype=act, device=device)
        for i in range(n_layers):
            p, q = f'encoder.layers.{i}.', f'L{i}.'
            a = p + 'attention.'
            t[q + 'attn.wqkv'] = W(torch.cat([sd[a + 'q_proj.weight'], sd[a + 'k_proj.weight'], sd[a + 'v_proj.weight']], 0))
            t[q + 'attn.bqkv'] = Bv(torch.cat([sd[a + 'q_proj.bias'], sd[a + 'k_proj.bias'], sd[a + 'v_proj.bias']], 0))
            t[q + 'attn.wo'] = W(sd[a + 'out_proj.weight'])
            t[q + 'attn.bo'] = Bv(sd[a + 'out_proj.bias'])
            t[q + 'ln.w'] = Fv(sd[p + 'layer_norm.weight'])
            t[q + 'ln.b'] = Fv(sd[p + 'layer_norm.bias'])
            t[q + 'ffn.w1'] = W(sd[p + 'feed_forward.intermediate_dense.weight'])
            t[q + 'ffn.b1'] = Bv(sd[p + 'feed_forward.intermediate_dense.bias'])
            t[q + 'ffn.w2'] = W(sd[p + 'feed_forward.output_dense.weight'])
            t[q + 'ffn.b2'] = Bv(sd[p + 'feed_forward.output_dense.bias'])
            t[q + 'final.ln.w'] = Fv(sd[p + 'final_layer_norm.weight'])
            t[q + 'final.ln.b'] = Fv(sd[p + 'final_layer_norm.bias'])
        self.tensors = t


class HubertEncoder(torch.nn.Module):
    """semantic_s as the reference builds it: waveform (processor output) -> int16 tokens [B, 1, T].

    precision 'bf16' = the reference's CUDA autocast numerics on tcgen05 tensor cores, 'fp32' = CUDA-core fp32 (the
    reference's CPU numerics; used for the 1e-4 check).  ``normalize=True`` applies the feature extractor's zero-mean /
    unit-variance step on the device first (the reference does it in its dataset transform, encoder.py:20-26)."""

    num_codebooks = 1
    max_rows_per_batch = 65536

    def __init__(self, config=None, device: str = 'cuda:0', quantize: bool = True,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, codebook: Optional[torch.Tensor] = None,
                 precision: str = 'bf16', n_layers: Optional[int] = None, seed: int = 0):
        super().__init__()
        from .weights import synthetic_codebook, synthetic_hubert_state_dict
        self.config = config if config is not None else HubertEncoderConfig()
        self.device = torch.device(device)
        L.require_device(self.device)
        self.lib = L.load()
        self.precision = precision
        self.output_layer = self.config.output_layer
        self.n_layers = self.output_layer if n_layers is None else n_layers
        if state_dict is None:
            state_dict = synthetic_hubert_state_dict(seed, max(self.n_layers, 1))
        if codebook is None:
            codebook = synthetic_codebook(self.config.codebook_size, HID, seed=4)
        self.codebook_size = int(codebook.shape[0])
        with torch.cuda.device(self.device):
            self.weights = HubertWeights(state_dict, self.n_layers, self.device, precision)
            self.codebook = codebook.to(self.device, torch.float32).contiguous()
            self.handle = self.lib.b2t_hubert_create(self.n_layers, self.codebook_size, L.PREC_BF16 if precision == 'bf16' else L.PREC_FP32)
            if not self.handle:
                raise L.B2TError('b2t_hubert_create failed: ' + self.lib.b2t_last_error().decode())
            for name, t in self.weights.tensors.items():
                L.check(self.lib.b2t_hubert_set_tensor(self.handle, name.encode(), t.data_ptr()), name)
            L.check(self.lib.b2t_hubert_set_tensor(self.handle, b'codebook', self.codebook.data_ptr()), 'codebook')
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.b2t_hubert_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def rows_for(self, padded_samples: int) -> int:
        return int(feat_lengths(padded_samples))

    def rows_for_tokens(self, n_tokens: int, padded_samples: int) -> int:
        return max(1, min(n_tokens, self.rows_for(padded_samples)))

    def encode_plan(self, wave: torch.Tensor, plan: HubertPlan, normalize: bool = False, tap_layer: int = -1,
                    want_feats: bool = False):
        """wave: flat fp32 device tensor -> tokens int16 [total_rows] (+ tapped hidden state [total_rows, 768] and/or the
        feature-encoder output [total_rows, 512])."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.is_contiguous()
        with torch.cuda.device(self.device):
            db = DeviceHubertBatch(plan, self.device)
            need = self.lib.b2t_hubert_workspace_bytes(self.handle, db.byref())
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(int(need * 1.05) + 1024, dtype=torch.uint8, device=self.device)
            tokens = torch.empty(plan.total_rows, dtype=torch.int16, device=self.device)
            tap = torch.empty(plan.total_rows, HID, dtype=torch.float32, device=self.device) if tap_layer >= 0 else None
            feats = torch.empty(plan.total_rows, CONV_DIM, dtype=torch.float32, device=self.device) if want_feats else None
            L.check(self.lib.b2t_hubert_encode(self.handle, wave.data_ptr(), db.byref(), self._ws.data_ptr(), self._ws.numel(),
                                               int(normalize), tokens.data_ptr(), tap_layer, L.ptr(tap), L.ptr(feats),
                                               L.stream_ptr()), 'b2t_hubert_encode')
            self.last_launches = self.lib.b2t_last_launch_count()
            self._keep = db
        return tokens, tap, feats

    def forward(self, input_batch: torch.Tensor, attention_mask: torch.Tensor, tap_layer: int = -1):
        """input_batch [B, L] (processor output, right zero-padded), attention_mask [B, L] 0/1 -> int16 [B, 1, T]."""
        assert input_batch.dim() == 2, "Input tensor must have shape [batch, time]"
        B, Lp = input_batch.shape
        wave = input_batch.to(self.device, torch.float32).contiguous()
        lengths = attention_mask.to(self.device).sum(dim=1).round().to(torch.int64).cpu().numpy()
        plan = plan_hubert(lengths, np.arange(B, dtype=np.int64) * Lp, Lp)
        tokens, tap, _ = self.encode_plan(wave.view(-1), plan, False, tap_layer)
        T = plan.total_rows // B
        out = tokens.view(B, 1, T)
        if tap_layer >= 0:
            return out, tap.view(B, T, HID)
        return out

    def encode_single(self, audio: torch.Tensor) -> torch.Tensor:
        """AudioToken.encode on one array [1, L]: processor transform over the whole array, mask of ones (reference
        core.py:187-196) -> int16 [1, 1, T]."""
        n = int(audio.shape[-1])
        toks = self.encode_packed([audio.reshape(-1)], n, None, normalize=True)[0]
        return toks.view(1, 1, -1)

    def encode_packed(self, clips: Sequence[torch.Tensor], padded_samples, rows: Optional[Sequence[int]] = None,
                      normalize: bool = True) -> List[torch.Tensor]:
        """clips: raw 1-D fp32 waveforms (any device) -> int16 [1, rows_i] device tensors.  Every clip is normalised on
        its own (the reference's batch reader normalises every streamed chunk, datasets.py:75-79)."""
        lengths = [int(c.numel()) for c in clips]
        offs = np.zeros(len(clips), dtype=np.int64)
        offs[1:] = np.cumsum(lengths)[:-1]
        wave = torch.cat([c.reshape(-1).to(torch.float32) for c in clips]).to(self.device)
        if rows is not None:
            tv = feat_lengths(lengths)
            rows = [max(int(r), int(t)) for r, t in zip(rows, tv)]
        plan = plan_hubert(lengths, offs, padded_samples, rows)
        tokens, _, _ = self.encode_plan(wave, plan, normalize)
        ro = plan.row_off
        return [tokens[ro[i]:ro[i + 1]].view(1, -1) for i in range(len(clips))]
