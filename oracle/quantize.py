"""Oracle: nearest-centroid / VQ / residual-VQ assignment.  TEST INFRASTRUCTURE ONLY.

Anchors (the quantiser libraries themselves are third-party and absent from /root/reference):
  * k-means assignment ...... reference audiotoken/encoder.py:100-101 (`torch.cdist` + `argmin`)
  * semantic VQ ............. reference audiotoken/encoder.py:147-161, 180 —
    `vector_quantize_pytorch.VectorQuantize` (unpinned, requirements.txt:10) in eval mode with
    a Euclidean codebook: indices = argmax_k -cdist(x, embed[0]); codebook tensor
    `_codebook.embed` [1, K, D] (scripts/clustering/cluster_tokens.py:316-320).
  * residual VQ ............. reference audiotoken/encoder.py:50-52 — `encodec` (unpinned,
    requirements.txt:5) `quantizer.encode(emb, frame_rate, bandwidth)`; algorithm mirrored by
    transformers models/encodec/modeling_encodec.py:364-384, 395-404, 416-438:
    n_q = floor(bw*1000 / (log2(1024) * frame_rate)); per stage
    idx = argmax -(|r|^2 - 2 r.E^T + |E|^2); r -= E[idx].

The exact oracle works in float64 and reports near-ties (relative top-2 margin below 1e-6)
so that the bit-exactness claim excludes rows that fp32 arithmetic cannot order.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch


def nearest_centroid(x: torch.Tensor, codebook: torch.Tensor, chunk: int = 8192,
                     tie_rel_margin: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [M, D], codebook [K, D] -> (indices int64 [M], near_tie bool [M]).

    Exact float64 squared distances; first index on ties (torch argmin semantics).
    """
    x = x.detach().cpu().to(torch.float64)
    c = codebook.detach().cpu().to(torch.float64)
    cn = (c * c).sum(1)
    idx = torch.empty(x.shape[0], dtype=torch.int64)
    tie = torch.empty(x.shape[0], dtype=torch.bool)
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        d = (xs * xs).sum(1, keepdim=True) - 2.0 * (xs @ c.t()) + cn.unsqueeze(0)
        if c.shape[0] > 1:
            top2, i2 = torch.topk(d, 2, dim=1, largest=False)
            idx[s:s + chunk] = torch.argmin(d, dim=1)
            tie[s:s + chunk] = (top2[:, 1] - top2[:, 0]) <= tie_rel_margin * top2[:, 1].abs().clamp(min=1e-30)
        else:
            idx[s:s + chunk] = 0
            tie[s:s + chunk] = False
    return idx, tie


def kmeans_assign_fp32(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """The reference's own expression (encoder.py:100-101), fp32 on CPU."""
    d = torch.cdist(x.float().unsqueeze(0), centroids.float().unsqueeze(0))[0]
    return torch.argmin(d, dim=-1)


def num_quantizers(bandwidth: float, frame_rate: int = 75, bins: int = 1024) -> int:
    bw_per_q = math.log2(bins) * frame_rate / 1000.0
    return int(max(1, math.floor(bandwidth / bw_per_q)))


def rvq_encode(emb: torch.Tensor, codebooks: torch.Tensor, n_q: int,
               dtype=torch.float64) -> torch.Tensor:
    """emb [M, D], codebooks [n_total, K, D] -> codes int64 [n_q, M] (residual VQ)."""
    r = emb.detach().cpu().to(dtype).clone()
    out = []
    for q in range(n_q):
        E = codebooks[q].detach().cpu().to(dtype)
        d = (r * r).sum(1, keepdim=True) - 2.0 * (r @ E.t()) + (E * E).sum(1).unsqueeze(0)
        i = torch.argmin(d, dim=1)          # == argmax(-d), first index on ties
        out.append(i)
        r = r - E[i]
    return torch.stack(out, 0)


def vq_ema_train_step(x: torch.Tensor, embed: torch.Tensor, embed_avg: torch.Tensor, cluster_size: torch.Tensor,
                      decay: float = 0.8, eps: float = 1e-5, commitment_weight: float = 1.0, dtype=torch.float64):
    """One training-mode forward of the reference's quantiser (scripts/clustering/cluster_tokens.py:142-147, 293-311:
    `VectorQuantize(dim, codebook_size, decay=0.8, commitment_weight=1)` called on a batch).  The class is third-party
    (`vector_quantize_pytorch`, unpinned, requirements.txt:10); this restates the published training forward of its
    Euclidean codebook (one head, ema_update, threshold_ema_dead_code = 0, codebook already initialised):

        idx = argmin cdist(x, embed);  quantize = embed[idx];  loss = mse(quantize, x) * commitment_weight
        cluster_size <- cluster_size*decay + onehot.sum(0)*(1-decay)
        embed_avg    <- embed_avg*decay + (onehot^T x)*(1-decay)
        embed        <- embed_avg / (laplace_smoothing(cluster_size, K, eps) * cluster_size.sum())[:, None]

    x [M, D]; returns (idx int64 [M], loss float, embed', embed_avg', cluster_size') in `dtype`.
    PARITY UNPINNED against the third-party package itself (absent offline); anchored on the call site only."""
    x = x.detach().cpu().to(dtype)
    e = embed.detach().cpu().to(dtype)
    K = e.shape[0]
    idx, _ = nearest_centroid(x, e)
    q = e[idx]
    loss = float(((q - x) ** 2).mean() * commitment_weight)
    n = torch.bincount(idx, minlength=K).to(dtype)
    s = torch.zeros_like(e).index_add_(0, idx, x)
    cs = cluster_size.detach().cpu().to(dtype) * decay + n * (1.0 - decay)
    avg = embed_avg.detach().cpu().to(dtype) * decay + s * (1.0 - decay)
    tot = cs.sum()
    smoothed = (cs + eps) / (tot + K * eps) * tot
    return idx, loss, avg / smoothed.unsqueeze(1), avg, cs
