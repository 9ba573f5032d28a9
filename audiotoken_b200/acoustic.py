"""Acoustic tokenizer: host side of the EnCodec 24 kHz encode path.

Mirror of the reference's ``AcousticEncoder`` (audiotoken/encoder.py:29-57): called as
``encoder(input_batch[B, L], attention_mask) -> int16 [B, n_q, ceil(L/320)]`` (the mask is ignored, as in the
reference), plus ``encode_packed`` for ragged batches.  Weight preparation (weight-norm folding, tap-major
reordering), batch planning and buffer ownership live here; all arithmetic is in libb200tok.so
(csrc/acoustic.cu).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import lib as L
from .configs import AcousticEncoderConfig
from .weights import SEANET_CONVS, synthetic_encodec_state_dict, weight_norm_weight

STRIDES = (2, 4, 5, 8)
HOP = 320
TILE = 64


@dataclass
class AcousticPlan:
    n_clips: int
    wave_off: np.ndarray          # int64 [n]
    true_len: np.ndarray          # int32 [n]
    lens: List[np.ndarray]        # 5 x int32 [n]
    offs: List[np.ndarray]        # 5 x int32 [n+1]
    tile_clip: List[np.ndarray]
    tile_t0: List[np.ndarray]
    order: np.ndarray             # int32 [n]
    active: np.ndarray            # int32 [t_max]
    rank: np.ndarray              # int32 [n]   inverse of order
    toff: np.ndarray              # int32 [t_max + 1]   prefix sums of active
    aligned320: bool              # every virtual length is a multiple of 320 samples (tensor-core path)

    @property
    def frames(self) -> np.ndarray:
        return self.lens[4]

    @property
    def total_frames(self) -> int:
        return int(self.offs[4][-1])


def plan_acoustic(true_lens: Sequence[int], wave_offsets: Sequence[int], virt_lens: Sequence[int],
                  tiles: bool = True) -> AcousticPlan:
    """true_lens[i] samples are read from the wave buffer, the clip behaves as a signal of virt_lens[i]
    samples (zeros after true_len) — the reference's right zero-padding (datasets.py:99-103).
    tiles=False skips the 64-row work lists only the fp32 kernels use."""
    n = len(true_lens)
    tl = np.asarray(true_lens, dtype=np.int64)
    vl = np.asarray(virt_lens, dtype=np.int64)
    if n == 0 or np.any(tl < 1) or np.any(vl < tl):
        raise ValueError('clips must be non-empty and virt_len >= true_len')
    lens = [vl]
    for s in STRIDES:
        lens.append(-(-lens[-1] // s))
    offs, tc, tt = [], [], []
    for l in range(5):
        o = np.zeros(n + 1, dtype=np.int64)
        o[1:] = np.cumsum(lens[l])
        if o[-1] >= 2 ** 31:
            raise ValueError('batch too large for int32 offsets')
        offs.append(o.astype(np.int32))
        cl, t0 = [np.zeros(0, dtype=np.int32)], [np.zeros(0, dtype=np.int32)]
        for i in range(n if tiles else 0):
            t = np.arange(0, lens[l][i], TILE, dtype=np.int32)
            cl.append(np.full(t.shape, i, dtype=np.int32))
            t0.append(t)
        tc.append(np.concatenate(cl))
        tt.append(np.concatenate(t0))
    order = np.argsort(-lens[4], kind='stable').astype(np.int32)
    t_max = int(lens[4].max())
    sorted_len = lens[4][order]
    active = (sorted_len[None, :] > np.arange(t_max)[:, None]).sum(axis=1).astype(np.int32)
    rank = np.empty(n, dtype=np.int32)
    rank[order] = np.arange(n, dtype=np.int32)
    toff = np.zeros(t_max + 1, dtype=np.int32)
    toff[1:] = np.cumsum(active)
    if not tiles and not np.all(vl % HOP == 0):
        raise ValueError('tiles=False requires clip lengths that are multiples of 320 samples')
    return AcousticPlan(n, np.asarray(wave_offsets, dtype=np.int64), tl.astype(np.int32),
                        [x.astype(np.int32) for x in lens], offs, tc, tt, order, active, rank, toff,
                        bool(np.all(vl % HOP == 0)))


class DeviceAcousticBatch:
    def __init__(self, plan: AcousticPlan, device):
        arrays = ([plan.true_len] + plan.lens + plan.offs + plan.tile_clip + plan.tile_t0 +
                  [plan.order, plan.rank, plan.toff])
        n64 = plan.wave_off.size
        total = 2 * n64 + sum(a.size for a in arrays)
        host = torch.empty(total, dtype=torch.int32, pin_memory=torch.cuda.is_available())
        hv = host.numpy()
        hv[:2 * n64] = plan.wave_off.view(np.int32)
        off, ptrs = 2 * n64, []
        for a in arrays:
            hv[off:off + a.size] = a
            ptrs.append(off)
            off += a.size
        self.host = host
        self.dev = host.to(device, non_blocking=True)
        base = self.dev.data_ptr()
        p = [base + 4 * o for o in ptrs]
        b = L.AcousticBatch()
        b.n_clips = plan.n_clips
        b.t_max = int(plan.active.size)
        for l in range(5):
            b.total[l] = int(plan.offs[l][-1])
            b.n_tiles[l] = int(plan.tile_clip[l].size)
            b.len[l] = p[1 + l]
            b.off[l] = p[6 + l]
            b.tile_clip[l] = p[11 + l]
            b.tile_t0[l] = p[16 + l]
        b.wave_off = base
        b.true_len = p[0]
        b.order = p[21]
        b.rank = p[22]
        b.toff = p[23]
        self.frames = np.ascontiguousarray(plan.lens[4], dtype=np.int32)
        b.frames_host = self.frames.ctypes.data
        b.aligned320 = int(plan.aligned320)
        self.c = b
        self.active = np.ascontiguousarray(plan.active)

    def byref(self):
        return C.byref(self.c)


class AcousticWeights:
    """HF EncodecModel-named state dict -> fp32 device tensors for csrc/acoustic.cu."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, n_q_total: int):
        t: Dict[str, torch.Tensor] = {}
        for i, (name, cin, cout, k, _s) in enumerate(SEANET_CONVS):
            w = weight_norm_weight(sd, name).float()                       # [cout, cin, k]
            w = w.permute(0, 2, 1).reshape(cout, k * cin)                  # tap-major: index = tap * cin + ci
            kpad = (k * cin + 15) // 16 * 16
            wp = torch.zeros(cout, kpad)
            wp[:, :k * cin] = w
            t[f'conv{i}.w'] = wp.to(device).contiguous()
            t[f'conv{i}.b'] = sd[name + '.bias'].float().to(device).contiguous()
        for layer in range(2):
            p = 'encoder.layers.13.lstm.'
            t[f'lstm{layer}.w_ih'] = sd[p + f'weight_ih_l{layer}'].float().to(device).contiguous()
            t[f'lstm{layer}.w_hh'] = sd[p + f'weight_hh_l{layer}'].float().to(device).contiguous()
            t[f'lstm{layer}.b'] = (sd[p + f'bias_ih_l{layer}'] + sd[p + f'bias_hh_l{layer}']).float().to(device).contiguous()
        cb = torch.stack([sd[f'quantizer.layers.{q}.codebook.embed'].float() for q in range(n_q_total)])
        hn = (0.5 * (cb.double() ** 2).sum(-1))
        t['rvq.codebooks'] = cb.to(device).contiguous()
        t['rvq.half_norm'] = hn.float().to(device).contiguous()
        t['rvq.cmax_half'] = hn.max(dim=1).values.float().to(device).contiguous()
        # tensor-core RVQ (csrc/rvq_tc.cu): error-compensated bf16 pair of every codebook, [n_q*1024, hi(128) | lo(128)]
        hi = cb.to(torch.bfloat16)
        lo = (cb - hi.float()).to(torch.bfloat16)
        t['rvq.c2'] = torch.cat([hi, lo], dim=-1).reshape(-1, 256).to(device).contiguous()
        t['rvq.stats'] = torch.zeros(2, dtype=torch.int32, device=device)
        # tensor-core path (csrc/seanet_tc.cu): bf16 weights, K padded to 64, tap-major; weight-norm is evaluated in
        # fp32 and then rounded, as autocast does
        def tapmajor(i):
            name, cin, cout, k, _s = SEANET_CONVS[i]
            return weight_norm_weight(sd, name).float().permute(0, 2, 1).reshape(cout, k * cin), sd[name + '.bias'].float()

        def pad64(w):
            kp = (w.shape[1] + 63) // 64 * 64
            out = torch.zeros(w.shape[0], kp)
            out[:, :w.shape[1]] = w
            return out.to(torch.bfloat16).to(device).contiguous()

        for l in range(4):
            w3, b3 = tapmajor(1 + 4 * l)
            wk1, bk1 = tapmajor(2 + 4 * l)
            wsc, bsc = tapmajor(3 + 4 * l)
            wd, bd = tapmajor(4 + 4 * l)
            t[f'tc.k3{l}.w'], t[f'tc.k3{l}.b'] = pad64(w3), b3.to(device).contiguous()
            t[f'tc.res{l}.w'] = pad64(torch.cat([wsc, wk1], dim=1))
            t[f'tc.res{l}.b'] = (bsc + bk1).to(device).contiguous()
            t[f'tc.down{l}.w'], t[f'tc.down{l}.b'] = pad64(wd), bd.to(device).contiguous()
        wf, bf = tapmajor(17)
        t['tc.final.w'], t['tc.final.b'] = pad64(wf), bf.to(device).contiguous()
        for layer in range(2):
            p = 'encoder.layers.13.lstm.'
            w = torch.cat([sd[p + f'weight_ih_l{layer}'].float(), sd[p + f'weight_hh_l{layer}'].float()], dim=1)
            bsum = (sd[p + f'bias_ih_l{layer}'] + sd[p + f'bias_hh_l{layer}']).float()
            # row 4*u + g  <-  gate g (i, f, g, o) of hidden unit u
            t[f'tc.lstm{layer}.w'] = w.view(4, 512, 1024).permute(1, 0, 2).reshape(2048, 1024).to(torch.bfloat16).to(device).contiguous()
            t[f'tc.lstm{layer}.b'] = bsum.view(4, 512).t().reshape(2048).to(device).contiguous()
        self.tensors = t


class AcousticEncoder(torch.nn.Module):
    """Same call shape as reference audiotoken/encoder.py:29-57.

    precision='bf16' (default): tcgen05 encoder, bf16 operands / fp32 accumulation — the reference's GPU numerics
    (autocast, encoder.py:45); batches whose clip lengths are not multiples of 320 samples run the fp32 kernels.
    precision='fp32': CUDA-core fp32 kernels — the reference's CPU numerics (BASELINE config 1).
    The residual VQ is exact (fp32 residuals, fp64-checked argmin) in both."""

    def __init__(self, config: Optional[AcousticEncoderConfig] = None, device: str = 'cuda:0',
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0, precision: str = 'bf16',
                 **_unused):
        super().__init__()
        assert precision in ('bf16', 'fp32')
        self.precision = precision
        # frames per ragged batch: the fp32 path keeps every level of the whole batch resident (~290 KB/frame);
        # the tensor path streams the strided front end in sub-batches and keeps only 75 Hz tensors (~4 KB/frame)
        self.max_rows_per_batch = 75 * 1200 if precision == 'fp32' else 75 * 40000
        self.config = config if config is not None else AcousticEncoderConfig()
        self.device = torch.device(device)
        L.require_device(self.device)
        self.lib = L.load()
        self.num_codebooks = self.config.num_codebooks           # floor(bw*1000 / (10*75))
        if state_dict is None:
            state_dict = synthetic_encodec_state_dict(seed)
        n_total = sum(1 for k in state_dict if k.endswith('codebook.embed'))
        assert self.num_codebooks <= n_total
        with torch.cuda.device(self.device):
            self.weights = AcousticWeights(state_dict, self.device, n_total)
            self.handle = self.lib.b2t_acoustic_create()
            for name, t in self.weights.tensors.items():
                L.check(self.lib.b2t_acoustic_set_tensor(self.handle, name.encode(), t.data_ptr()), name)
        self._ws: Optional[torch.Tensor] = None
        self.last_launches = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.b2t_acoustic_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def rows_for(self, padded_samples: int) -> int:
        return -(-padded_samples // HOP)

    def rows_for_tokens(self, n_tokens: int, padded_samples: int) -> int:
        return max(1, min(n_tokens, self.rows_for(padded_samples)))

    def encode_plan(self, wave: torch.Tensor, plan: AcousticPlan, want_emb: bool = False):
        """-> codes int16 [n_q, total_frames] (and fp32 embeddings [total_frames, 128])."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.is_contiguous()
        with torch.cuda.device(self.device):
            db = DeviceAcousticBatch(plan, self.device)
            prec = L.PREC_BF16 if (self.precision == 'bf16' and plan.aligned320) else L.PREC_FP32
            if prec == L.PREC_FP32 and plan.tile_clip[0].size == 0:
                raise ValueError('plan was built with tiles=False but the fp32 kernels are needed')
            self.last_precision = 'bf16' if prec == L.PREC_BF16 else 'fp32'
            need = self.lib.b2t_acoustic_workspace_bytes(db.byref(), prec)
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(int(need * 1.05) + 1024, dtype=torch.uint8, device=self.device)
            codes = torch.empty(self.num_codebooks, plan.total_frames, dtype=torch.int16, device=self.device)
            emb = torch.empty(plan.total_frames, 128, device=self.device) if want_emb else None
            L.check(self.lib.b2t_acoustic_encode(self.handle, wave.data_ptr(), db.byref(), self.num_codebooks, prec,
                                                 self._ws.data_ptr(), self._ws.numel(), codes.data_ptr(), L.ptr(emb),
                                                 db.active.ctypes.data, L.stream_ptr()), 'b2t_acoustic_encode')
            self.last_launches = self.lib.b2t_last_launch_count()
            self._keep = db
        return codes, emb

    def rvq_encode(self, emb: torch.Tensor, impl: int = L.IMPL_AUTO, n_q: Optional[int] = None) -> torch.Tensor:
        """emb fp32 [rows, 128] (device) -> codes int16 [n_q, rows]; the quantiser stage alone."""
        assert emb.is_cuda and emb.dtype == torch.float32 and emb.is_contiguous() and emb.shape[1] == 128
        n_q = self.num_codebooks if n_q is None else n_q
        codes = torch.empty(n_q, emb.shape[0], dtype=torch.int16, device=emb.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.b2t_rvq_encode(self.handle, emb.data_ptr(), emb.shape[0], n_q, impl, codes.data_ptr(),
                                            L.stream_ptr()), 'b2t_rvq_encode')
        return codes

    def rvq_stats(self):
        """(fp64 re-scores, exhaustive re-scans) taken by the tensor RVQ kernel since the last call."""
        s = self.weights.tensors['rvq.stats']
        out = tuple(int(v) for v in s.cpu())
        s.zero_()
        return out

    def forward(self, input_batch: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, want_emb: bool = False):
        """input_batch [B, L] fp32 -> int16 [B, n_q, ceil(L/320)]; attention_mask is ignored (encoder.py:44)."""
        assert input_batch.dim() == 2
        B, Lp = input_batch.shape
        wave = input_batch.to(self.device, torch.float32).contiguous()
        plan = plan_acoustic([Lp] * B, np.arange(B, dtype=np.int64) * Lp, [Lp] * B,
                             tiles=not (self.precision == 'bf16' and Lp % HOP == 0))
        codes, emb = self.encode_plan(wave.view(-1), plan, want_emb)
        T = plan.total_frames // B
        out = codes.view(self.num_codebooks, B, T).transpose(0, 1).contiguous()
        if want_emb:
            return out, emb.view(B, T, 128).transpose(1, 2)
        return out

    def encode_packed(self, clips: Sequence[torch.Tensor], padded_samples, rows: Optional[Sequence[int]] = None
                      ) -> List[torch.Tensor]:
        """Ragged batch -> list of int16 [n_q, rows_i]; clip i is zero-extended to rows_i * 320 samples."""
        lengths = [int(c.numel()) for c in clips]
        if rows is None:
            rows = [-(-n // HOP) for n in lengths]
        virt = [max(r * HOP, n) for r, n in zip(rows, lengths)]
        offs = np.zeros(len(clips), dtype=np.int64)
        offs[1:] = np.cumsum(lengths)[:-1]
        wave = torch.cat([c.reshape(-1).to(torch.float32) for c in clips]).to(self.device)
        plan = plan_acoustic(lengths, offs, virt, tiles=self.precision != 'bf16')
        codes, _ = self.encode_plan(wave, plan)
        fo = plan.offs[4]
        return [codes[:, fo[i]:fo[i] + rows[i]] for i in range(len(clips))]


class AcousticDecoder(torch.nn.Module):
    """Mirror of the reference's ``AcousticDecoder`` (audiotoken/decoder.py:50-76): ``decoder(tokens[B, K, T]) ->
    fp32 [1, B * T * 320]`` = ``model.decoder(model.quantizer.decode(tokens))`` flattened, on the device.
    fp32 CUDA-core kernels (csrc/acoustic.cu, b2t_acoustic_decode); no CPU fallback."""

    def __init__(self, config=None, device: str = 'cuda:0', state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 seed: int = 0, **_unused):
        super().__init__()
        from .weights import SEANET_DEC_CONVS
        self.device = torch.device(device)
        L.require_device(self.device)
        self.lib = L.load()
        if state_dict is None:
            state_dict = synthetic_encodec_state_dict(seed)
        sd = state_dict
        n_total = sum(1 for k in sd if k.endswith('codebook.embed'))
        t: Dict[str, torch.Tensor] = {}
        ci, ti = 0, 0
        for name, cin, cout, k, s, transposed in SEANET_DEC_CONVS:
            w = weight_norm_weight(sd, name).float()
            if transposed:                                        # [C_in, C_out, 2s] -> phase-major [s, C_out, 2*C_in]
                assert k == 2 * s
                wt = torch.cat([w[:, :, :s].permute(2, 1, 0), w[:, :, s:].permute(2, 1, 0)], dim=2)
                t[f'dec.convt{ti}.w'] = wt.contiguous().to(self.device)
                t[f'dec.convt{ti}.b'] = sd[name + '.bias'].float().to(self.device).contiguous()
                ti += 1
            else:                                                 # [C_out, C_in, k] -> tap-major, K padded to 16
                wm = w.permute(0, 2, 1).reshape(cout, k * cin)
                kpad = (k * cin + 15) // 16 * 16
                wp = torch.zeros(cout, kpad)
                wp[:, :k * cin] = wm
                t[f'dec.conv{ci}.w'] = wp.to(self.device).contiguous()
                t[f'dec.conv{ci}.b'] = sd[name + '.bias'].float().to(self.device).contiguous()
                ci += 1
        p = 'decoder.layers.1.lstm.'
        for layer in range(2):
            t[f'dec.lstm{layer}.w_ih'] = sd[p + f'weight_ih_l{layer}'].float().to(self.device).contiguous()
            t[f'dec.lstm{layer}.w_hh'] = sd[p + f'weight_hh_l{layer}'].float().to(self.device).contiguous()
            t[f'dec.lstm{layer}.b'] = (sd[p + f'bias_ih_l{layer}'] + sd[p + f'bias_hh_l{layer}']).float().to(self.device).contiguous()
        t['rvq.codebooks'] = torch.stack([sd[f'quantizer.layers.{q}.codebook.embed'].float() for q in range(n_total)]).to(self.device).contiguous()
        self.tensors = t
        with torch.cuda.device(self.device):
            self.handle = self.lib.b2t_acoustic_create()
            for name, ten in t.items():
                L.check(self.lib.b2t_acoustic_set_tensor(self.handle, name.encode(), ten.data_ptr()), name)
        self._ws: Optional[torch.Tensor] = None
        self.last_launches = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.b2t_acoustic_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def decode_packed(self, codes: torch.Tensor, frames: Sequence[int]) -> torch.Tensor:
        """codes int16 [n_q, sum(frames)] (packed clips) -> fp32 [sum(frames) * 320] on the device."""
        assert codes.is_cuda and codes.dtype == torch.int16 and codes.is_contiguous()
        n_q = codes.shape[0]
        lens = [int(f) * HOP for f in frames]
        offs = np.zeros(len(lens), dtype=np.int64)
        offs[1:] = np.cumsum(lens)[:-1]
        plan = plan_acoustic(lens, offs, lens)
        assert plan.total_frames == codes.shape[1]
        with torch.cuda.device(self.device):
            db = DeviceAcousticBatch(plan, self.device)
            need = self.lib.b2t_acoustic_decode_workspace_bytes(db.byref())
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(int(need) + 1024, dtype=torch.uint8, device=self.device)
            wave = torch.empty(int(plan.offs[0][-1]), dtype=torch.float32, device=self.device)
            L.check(self.lib.b2t_acoustic_decode(self.handle, codes.data_ptr(), db.byref(), n_q, self._ws.data_ptr(),
                                                 self._ws.numel(), wave.data_ptr(), db.active.ctypes.data, L.stream_ptr()),
                    'b2t_acoustic_decode')
            self.last_launches = self.lib.b2t_last_launch_count()
            self._keep = db
        return wave

    def forward(self, input_batch: torch.Tensor) -> torch.Tensor:
        """tokens [B, K, T] (any integer dtype) -> fp32 [1, B * T * 320]  (reference decoder.py:62-76)."""
        assert input_batch.dim() == 3
        B, K, T = input_batch.shape
        codes = input_batch.to(self.device).to(torch.int16).permute(1, 0, 2).reshape(K, B * T).contiguous()
        return self.decode_packed(codes, [T] * B).unsqueeze(0)
