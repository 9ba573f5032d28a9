"""Host-side batch planner: ragged clips -> the packed b2t_batch descriptor.

The reference pads every segment to ``chunk_size`` seconds and carries a sample-level mask
(audiotoken/datasets.py:88-105); the frame/row validity it derives from that mask
(audiotoken/processors.py:80-115, 192-207, 244-259) only depends on each clip's length and on
the parity of the padded frame count.  This module computes those integers directly:

  n_valid   = 1 + floor((len - 400) / 160)            frames fully inside the clip
  n_total   = 1 + floor((L_pad - 400) / 160)          frames of the (virtual) padded batch
  stack     = min(n_valid, 2 * floor(n_total / 2))    an odd trailing frame is dropped (:244-250)
  valid_rows= ceil(stack / 2)                         attention_mask = validity of sub-frame 0 (:204)
  T         = roundup(floor(n_total / 2), pad_to_multiple_of)   rows the reference returns

Pure numpy; tested on CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

FRAME, HOP = 400, 160
QTILE, CTILE, QTILE128 = 64, 64, 128


def n_frames(num_samples: int) -> int:
    return 1 + (num_samples - FRAME) // HOP if num_samples >= FRAME else 0


def padded_rows(padded_samples: int, pad_to_multiple_of: int = 2) -> int:
    """T of the reference for a batch padded to `padded_samples` (processors.py:244-259, 195-197)."""
    t = n_frames(padded_samples) // 2
    if pad_to_multiple_of > 0 and t % pad_to_multiple_of:
        t += pad_to_multiple_of - t % pad_to_multiple_of
    return t


def length_tokens(num_samples: int, sample_rate: int, token_rate: int) -> int:
    """reference configs.py:213-218"""
    return math.ceil(num_samples / sample_rate * token_rate)


@dataclass
class SemanticPlan:
    """Integer arrays of one packed batch (host side)."""
    n_clips: int
    wave_off: np.ndarray      # int64 [n]
    frame_off: np.ndarray     # int32 [n+1]
    stack_frames: np.ndarray  # int32 [n]
    row_off: np.ndarray       # int32 [n+1]
    valid_rows: np.ndarray    # int32 [n]
    qtile_clip: np.ndarray
    qtile_q0: np.ndarray
    ctile_clip: np.ndarray
    ctile_t0: np.ndarray
    qtile128_clip: np.ndarray
    qtile128_q0: np.ndarray

    @property
    def total_rows(self) -> int:
        return int(self.row_off[-1])

    @property
    def total_frames(self) -> int:
        return int(self.frame_off[-1])

    @property
    def rows(self) -> np.ndarray:
        return np.diff(self.row_off)


def attention_work_lists(rows_arr, valid_rows):
    """(clip, first query row) of every 64- and 128-query tile, heaviest (most keys) clips first so that the tail of the
    grid is made of short clips."""
    order = np.argsort(-np.asarray(valid_rows), kind='stable')
    qc, qq, qc8, qq8 = [], [], [], []
    for i in order:
        q0 = np.arange(0, rows_arr[i], QTILE, dtype=np.int32)
        qc.append(np.full(q0.shape, i, dtype=np.int32))
        qq.append(q0)
        q8 = np.arange(0, rows_arr[i], QTILE128, dtype=np.int32)
        qc8.append(np.full(q8.shape, i, dtype=np.int32))
        qq8.append(q8)
    return np.concatenate(qc), np.concatenate(qq), np.concatenate(qc8), np.concatenate(qq8)


def plan_semantic(lengths: Sequence[int], wave_offsets: Sequence[int], padded_samples: Sequence[int] | int,
                  rows: Optional[Sequence[int]] = None, pad_to_multiple_of: int = 2) -> SemanticPlan:
    """Plan a packed batch.

    lengths[i]        valid samples of clip i (>= 400)
    wave_offsets[i]   first sample of clip i in the wave buffer
    padded_samples    length (samples) the reference would have padded clip i to (int or per clip)
    rows[i]           token rows to compute for clip i (default: all T rows the reference returns)
    """
    n = len(lengths)
    lengths = np.asarray(lengths, dtype=np.int64)
    if np.any(lengths < FRAME):
        raise ValueError(f'every clip needs at least {FRAME} samples (got min {int(lengths.min())})')
    pad = np.broadcast_to(np.asarray(padded_samples, dtype=np.int64), (n,))
    if np.any(pad < lengths):
        raise ValueError('padded_samples must be >= the clip length')
    n_valid = 1 + (lengths - FRAME) // HOP
    n_total = 1 + (pad - FRAME) // HOP
    stack = np.minimum(n_valid, 2 * (n_total // 2))
    valid_rows = (stack + 1) // 2
    t_full = n_total // 2
    if pad_to_multiple_of > 0:
        t_full = (t_full + pad_to_multiple_of - 1) // pad_to_multiple_of * pad_to_multiple_of
    if rows is None:
        rows_arr = t_full.copy()
    else:
        rows_arr = np.asarray(rows, dtype=np.int64)
        if np.any(rows_arr > t_full) or np.any(rows_arr < 1):
            raise ValueError('rows[i] must be in [1, T_i]')
    if np.any(valid_rows < 1):
        raise ValueError('a clip has no valid frame')
    if np.any(rows_arr < valid_rows):
        # the attention kernels take valid_rows[i] keys from clip i's packed rows: fewer rows than keys would read the
        # next clip's rows (and differ from the reference, whose keys always cover the whole clip)
        i = int(np.argmax(rows_arr < valid_rows))
        raise ValueError(f'rows[{i}] = {int(rows_arr[i])} is smaller than the {int(valid_rows[i])} valid token rows of the clip')
    frame_off = np.zeros(n + 1, dtype=np.int64)
    frame_off[1:] = np.cumsum(n_valid)
    row_off = np.zeros(n + 1, dtype=np.int64)
    row_off[1:] = np.cumsum(rows_arr)
    if row_off[-1] >= 2 ** 31 or frame_off[-1] >= 2 ** 31:
        raise ValueError('batch too large for int32 offsets')
    qc, qq, qc8, qq8 = attention_work_lists(rows_arr, valid_rows)
    cc, ct = [], []
    for i in range(n):
        t0 = np.arange(0, rows_arr[i], CTILE, dtype=np.int32)
        cc.append(np.full(t0.shape, i, dtype=np.int32))
        ct.append(t0)
    return SemanticPlan(
        n_clips=n,
        wave_off=np.asarray(wave_offsets, dtype=np.int64),
        frame_off=frame_off.astype(np.int32), stack_frames=stack.astype(np.int32),
        row_off=row_off.astype(np.int32), valid_rows=valid_rows.astype(np.int32),
        qtile_clip=qc, qtile_q0=qq,
        ctile_clip=np.concatenate(cc), ctile_t0=np.concatenate(ct),
        qtile128_clip=qc8, qtile128_q0=qq8)


class DeviceBatch:
    """A SemanticPlan uploaded to the device + the ctypes b2t_batch that points into it."""

    def __init__(self, plan: SemanticPlan, device):
        import torch
        from . import lib as L
        self.plan = plan
        i32 = [plan.frame_off, plan.stack_frames, plan.row_off, plan.valid_rows, plan.qtile_clip,
               plan.qtile_q0, plan.ctile_clip, plan.ctile_t0, plan.qtile128_clip, plan.qtile128_q0]
        # one pinned staging buffer, one H2D copy; int64 wave offsets first (8-byte aligned)
        n64 = plan.wave_off.size
        sizes = [a.size for a in i32]
        total = 2 * n64 + sum(sizes)
        host = torch.empty(total, dtype=torch.int32, pin_memory=torch.cuda.is_available())
        hv = host.numpy()
        hv[:2 * n64] = plan.wave_off.view(np.int32)
        off = 2 * n64
        offs = []
        for a in i32:
            hv[off:off + a.size] = a
            offs.append(off)
            off += a.size
        self.host = host
        self.dev = host.to(device, non_blocking=True)
        base = self.dev.data_ptr()
        p = [base + 4 * o for o in offs]
        b = L.Batch()
        b.n_clips = plan.n_clips
        b.total_frames = plan.total_frames
        b.total_rows = plan.total_rows
        b.n_qtiles = int(plan.qtile_clip.size)
        b.n_ctiles = int(plan.ctile_clip.size)
        b.max_rows = int(plan.rows.max()) if plan.n_clips else 0
        b.wave_off = base
        (b.frame_off, b.stack_frames, b.row_off, b.valid_rows, b.qtile_clip, b.qtile_q0,
         b.ctile_clip, b.ctile_t0, b.qtile128_clip, b.qtile128_q0) = p
        b.n_qtiles128 = int(plan.qtile128_clip.size)
        self.c = b

    @property
    def h2d_bytes(self) -> int:
        return self.host.numel() * 4

    def byref(self):
        return C.byref(self.c)


def bucket_by_rows(rows: Sequence[int], max_rows_per_batch: int) -> List[List[int]]:
    """Length-bucketed batching: sort clips by length (descending) and cut batches at a row budget.

    Because the layout is packed there is no padding waste; bucketing only keeps the attention
    work items of a batch similar in size and bounds the workspace."""
    order = sorted(range(len(rows)), key=lambda i: -rows[i])
    batches, cur, tot = [], [], 0
    for i in order:
        if cur and tot + rows[i] > max_rows_per_batch:
            batches.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += rows[i]
    if cur:
        batches.append(cur)
    return batches
