"""Codebook training: the EMA k-means step behind the reference's clustering script.

Mirrors how ``scripts/clustering/cluster_tokens.py`` uses ``vector_quantize_pytorch.VectorQuantize``:
``get_vq_model`` (:142-166) builds ``VectorQuantize(dim, codebook_size, decay=0.8, commitment_weight=1)`` and loads a
checkpoint; the training loop (:293-311) calls it on every batch of LayerNormed embeddings
(``_, indices, commit_loss = quantizer(x)``) and saves ``state_dict()`` (:316-320) with the keys
``_codebook.embed [1,K,D]``, ``_codebook.embed_avg [1,K,D]``, ``_codebook.cluster_size [1,K]``, ``_codebook.initted``.

Both halves run in libb200tok.so: the assignment is ``b2t_vq_argmin`` (exact, tcgen05 fast pass), the update is
``b2t_vq_ema_update`` (deterministic segmented sums).  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import lib as L


class CodebookTrainer:
    """Drop-in for the training-mode ``VectorQuantize`` of the clustering script (Euclidean codebook, one head, EMA
    update, no dead-code expiry, no k-means initialisation: the script always starts from a checkpoint or data-derived
    rows, cluster_tokens.py:150-158)."""

    def __init__(self, dim: int, codebook_size: int, decay: float = 0.8, commitment_weight: float = 1.0,
                 eps: float = 1e-5, device='cuda:0', return_quantized: bool = False):
        self.device = torch.device(device)
        L.require_device(self.device)
        self.lib = L.load()
        self.dim, self.codebook_size = int(dim), int(codebook_size)
        self.decay, self.commitment_weight, self.eps = float(decay), float(commitment_weight), float(eps)
        self.return_quantized = return_quantized
        self.embed = torch.zeros(codebook_size, dim, device=self.device)
        self.embed_avg = torch.zeros(codebook_size, dim, device=self.device)
        self.cluster_size = torch.zeros(codebook_size, device=self.device)
        self.training = True
        self._ws_vq: Optional[torch.Tensor] = None
        self._ws_ema: Optional[torch.Tensor] = None

    # ---- checkpoint format of the reference (cluster_tokens.py:316-320, utils.py:331-339) ----
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {'_codebook.initted': torch.tensor([True]),
                '_codebook.cluster_size': self.cluster_size.detach().cpu().unsqueeze(0).clone(),
                '_codebook.embed_avg': self.embed_avg.detach().cpu().unsqueeze(0).clone(),
                '_codebook.embed': self.embed.detach().cpu().unsqueeze(0).clone()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        emb = sd['_codebook.embed']
        if tuple(emb.shape) != (1, self.codebook_size, self.dim):
            raise ValueError(f'_codebook.embed has shape {tuple(emb.shape)}, expected {(1, self.codebook_size, self.dim)}')
        self.embed.copy_(emb[0].float())
        self.embed_avg.copy_(sd.get('_codebook.embed_avg', emb)[0].float())
        cs = sd.get('_codebook.cluster_size')
        self.cluster_size.copy_(cs[0].float() if cs is not None else torch.zeros(self.codebook_size))

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _workspace(self, name: str, need: int) -> torch.Tensor:
        ws = getattr(self, name)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            setattr(self, name, ws)
        return ws

    def __call__(self, x: torch.Tensor) -> Tuple[Optional[torch.Tensor], torch.Tensor, torch.Tensor]:
        """x [..., D] fp32 on the device -> (quantized or None, indices int64 [...], commit_loss [1]).
        In training mode the codebook state is updated in place (one EMA step over all rows of x)."""
        if x.device != self.device or x.dtype != torch.float32:
            raise ValueError('CodebookTrainer expects a float32 tensor on its device')
        if x.shape[-1] != self.dim:
            raise ValueError(f'last dimension {x.shape[-1]} != dim {self.dim}')
        lead = x.shape[:-1]
        flat = x.reshape(-1, self.dim)
        if flat.stride(-1) != 1 or flat.stride(0) % 4 or flat.data_ptr() % 16:
            flat = flat.contiguous()
        M, D, K = flat.shape[0], self.dim, self.codebook_size
        lib, st = self.lib, L.stream_ptr()
        ws = self._workspace('_ws_vq', lib.b2t_vq_workspace_bytes(M, D, K))
        i16 = torch.empty(M, dtype=torch.int16, device=self.device)
        i32 = torch.empty(M, dtype=torch.int32, device=self.device)
        L.check(lib.b2t_vq_argmin(flat.data_ptr(), flat.stride(0), M, D, self.embed.data_ptr(), None, K, 0, L.IMPL_AUTO,
                                  i16.data_ptr(), i32.data_ptr(), ws.data_ptr(), ws.numel(), st), 'vq_argmin')
        loss = torch.zeros(1, device=self.device)
        quant = torch.empty(M, D, device=self.device) if self.return_quantized else None
        if self.training:
            we = self._workspace('_ws_ema', lib.b2t_vq_ema_workspace_bytes(M, D, K))
            L.check(lib.b2t_vq_ema_update(flat.data_ptr(), flat.stride(0), M, D, i32.data_ptr(), self.embed.data_ptr(),
                                          self.embed_avg.data_ptr(), self.cluster_size.data_ptr(), K, self.decay,
                                          self.eps, self.commitment_weight, loss.data_ptr(),
                                          quant.data_ptr() if quant is not None else None, we.data_ptr(), we.numel(),
                                          st), 'vq_ema_update')
        elif quant is not None:
            raise NotImplementedError('return_quantized is only provided by the training step')
        return (quant.reshape(*lead, D) if quant is not None else None), i32.long().reshape(lead), loss



def iter_embeddings(encoder, files, chunk_size: int = 30, max_rows: int = 65536, layer: Optional[int] = None,
                    padded_rows: bool = False):
    """LayerNormed hidden states of `layer` (default: the encoder's output layer) for a list of audio files, one fp32
    device tensor [rows, 1024] per ragged batch of at most `max_rows` token rows: the feature stream the clustering script
    trains on (``get_features_batch``, cluster_tokens.py:83-136: affine-free LayerNorm of ``encoded_audio[l]``, reshaped to
    ``[B*T, D]``).  The reference pads every segment to ``chunk_size`` seconds and trains on the padded rows as well;
    `padded_rows=True` reproduces that (T rows per segment), the default feeds only the rows that become tokens.
    `encoder` is a Wav2VecBertEncoder; files are read, decoded and resampled on its device (io.read_audio_chunks)."""
    from . import io as aio
    from .packing import bucket_by_rows, length_tokens, plan_semantic
    sr, rate = encoder.config.model_sample_rate, encoder.config.model_token_rate
    layer = encoder.n_layers if layer is None else int(layer)
    pad = int(chunk_size * sr)
    pend, pend_rows = [], []                       # segments waiting for a batch

    def flush():
        for idx in bucket_by_rows(pend_rows, max_rows):
            clips = [pend[i] for i in idx]
            lengths = [int(c.numel()) for c in clips]
            offs = [0]
            for n in lengths[:-1]:
                offs.append(offs[-1] + n)
            plan = plan_semantic(lengths, offs, pad, None if padded_rows else [pend_rows[i] for i in idx])
            wave = torch.cat([c.reshape(-1) for c in clips]).to(encoder.device, torch.float32)
            _, hidden = encoder.encode_plan(wave, plan, tap_layer=layer)
            yield torch.nn.functional.layer_norm(hidden, (hidden.shape[-1],))
        pend.clear()
        pend_rows.clear()

    for path in files:
        for wave in aio.read_audio_chunks(path, sr, chunk_size, encoder.device):
            for seg in aio.iter_segments(wave, str(path), sr, rate, chunk_size):
                n = int(seg.wave.numel())
                pend.append(seg.wave)
                pend_rows.append(encoder.rows_for(pad) if padded_rows
                                 else encoder.rows_for_tokens(length_tokens(n, sr, rate), pad))
        if sum(pend_rows) >= max_rows:
            yield from flush()
    if pend:
        yield from flush()


def train_codebook(trainer: CodebookTrainer, batches, outdir: Optional[str] = None, save_freq: int = 10, layer: int = 19,
                   log=None) -> Dict[str, float]:
    """The clustering script's training loop (cluster_tokens.py:293-320): one EMA step of `trainer` per batch of
    embeddings (``_, indices, commit_loss = quantizer(x)``), the commitment loss and the share of active codes reported per
    batch, a checkpoint ``quantizer__L{layer}_C{K}_ckpt{idx}.pkl`` (``torch.save(state_dict())``) every `save_freq` batches.
    `batches`: any iterable of fp32 device tensors [rows, D], e.g. iter_embeddings(...)."""
    import os
    trainer.train()
    stats = {'batches': 0, 'rows': 0, 'commit_loss': float('nan'), 'active_fraction': float('nan')}
    for idx, x in enumerate(batches):
        _, indices, loss = trainer(x)
        stats['batches'] += 1
        stats['rows'] += int(x.shape[0])
        stats['commit_loss'] = float(loss.item())
        stats['active_fraction'] = indices.unique().numel() / trainer.codebook_size
        if log is not None:
            log(f"batch {idx}: rows {int(x.shape[0])}, commitment loss {stats['commit_loss']:.3f}, "
                f"active {100 * stats['active_fraction']:.3f} %")
        if outdir is not None and idx % save_freq == 0:
            os.makedirs(outdir, exist_ok=True)
            torch.save(trainer.state_dict(), os.path.join(outdir, f'quantizer__L{layer}_C{trainer.codebook_size}_ckpt{idx}.pkl'))
    return stats
