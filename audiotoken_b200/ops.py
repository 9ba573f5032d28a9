"""Thin torch-tensor wrappers over the individual C entry points (one function per kernel class).

Used by the parity tests and by anyone who wants a single stage (e.g. the quantiser microbench).
Inputs must already live on the sm_100 device; nothing here falls back to PyTorch maths.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import lib as L
from .fbank_tables import DeviceFbankTables
from .packing import DeviceBatch, SemanticPlan

_ACT = {L.PREC_BF16: torch.bfloat16, L.PREC_FP32: torch.float32}


def _prec(precision: str) -> int:
    return {'bf16': L.PREC_BF16, 'fp32': L.PREC_FP32}[precision]


def fbank_features(wave: torch.Tensor, plan: SemanticPlan, ln_w: torch.Tensor, ln_b: torch.Tensor,
                   precision: str = 'fp32', mel_bf16: bool = False):
    """-> (logmel [F,80], features [M,160] fp32 pre-LN, ln_out [M,160] act, row_valid [M] uint8)"""
    lib = L.load()
    dev = wave.device
    L.require_device(dev)
    db = DeviceBatch(plan, dev)
    tabs = DeviceFbankTables(dev)
    F_, M, n = plan.total_frames, plan.total_rows, plan.n_clips
    logmel = torch.empty(F_, 80, device=dev)
    mean = torch.empty(n, 80, device=dev)
    std = torch.empty(n, 80, device=dev)
    feats = torch.empty(M, 160, device=dev)
    out = torch.empty(M, 160, device=dev, dtype=_ACT[_prec(precision)])
    valid = torch.empty(M, device=dev, dtype=torch.uint8)
    s = L.stream_ptr()
    L.check(lib.b2t_fbank_logmel(wave.data_ptr(), db.byref(), tabs.byref(), logmel.data_ptr(), int(mel_bf16), s), 'fbank_logmel')
    L.check(lib.b2t_fbank_stats(logmel.data_ptr(), db.byref(), mean.data_ptr(), std.data_ptr(), s), 'fbank_stats')
    L.check(lib.b2t_fbank_stack_ln(logmel.data_ptr(), mean.data_ptr(), std.data_ptr(), db.byref(), ln_w.data_ptr(),
                                   ln_b.data_ptr(), out.data_ptr(), feats.data_ptr(), valid.data_ptr(),
                                   _prec(precision), s), 'fbank_stack_ln')
    torch.cuda.synchronize(dev)
    return logmel, feats, out, valid


def layernorm(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor],
              row_valid: Optional[torch.Tensor] = None, out_precision: str = 'fp32') -> torch.Tensor:
    lib = L.load()
    L.require_device(x.device)
    out = torch.empty(x.shape, device=x.device, dtype=_ACT[_prec(out_precision)])
    L.check(lib.b2t_layernorm(x.data_ptr(), L.ptr(w), L.ptr(b), L.ptr(row_valid), out.data_ptr(), x.shape[0],
                              x.shape[1], _prec(out_precision), L.stream_ptr()), 'layernorm')
    return out


def gemm(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int, precision: str,
         impl: int = L.IMPL_AUTO, resid: Optional[torch.Tensor] = None, row_valid: Optional[torch.Tensor] = None,
         alpha: float = 1.0, round_resid: bool = False) -> torch.Tensor:
    """Returns `out` (BIAS / SWISH / GLU) or the updated `resid` (RESID / BIAS_MASK)."""
    lib = L.load()
    L.require_device(A.device)
    M, K = A.shape
    N = W.shape[0]
    p = _prec(precision)
    g = L.GemmArgs()
    g.A, g.lda, g.W, g.bias = A.data_ptr(), A.stride(0), W.data_ptr(), L.ptr(bias)
    out = None
    if epilogue in (L.EPI_BIAS, L.EPI_BIAS_SWISH):
        out = torch.empty(M, N, device=A.device, dtype=_ACT[p])
    elif epilogue == L.EPI_GLU:
        out = torch.empty(M, N // 2, device=A.device, dtype=_ACT[p])
    elif epilogue == L.EPI_BIAS_MASK and resid is None:
        resid = torch.empty(M, N, device=A.device, dtype=torch.float32)
    g.out, g.ldo = (out.data_ptr(), out.stride(0)) if out is not None else (None, 0)
    g.resid, g.row_valid = L.ptr(resid), L.ptr(row_valid)
    g.M, g.N, g.K, g.epilogue, g.alpha = M, N, K, epilogue, alpha
    g.round_resid_bf16, g.precision, g.impl = int(round_resid), p, impl
    L.check(lib.b2t_gemm(C.byref(g), L.stream_ptr()), 'gemm')
    return out if out is not None else resid


def relkey_attention(qkv: torch.Tensor, dist_emb: torch.Tensor, plan: SemanticPlan, precision: str,
                     impl: int = L.IMPL_AUTO) -> torch.Tensor:
    lib = L.load()
    L.require_device(qkv.device)
    db = DeviceBatch(plan, qkv.device)
    out = torch.empty(qkv.shape[0], 1024, device=qkv.device, dtype=qkv.dtype)
    L.check(lib.b2t_relkey_attention(qkv.data_ptr(), dist_emb.data_ptr(), db.byref(), out.data_ptr(),
                                     _prec(precision), impl, L.stream_ptr()), 'relkey_attention')
    torch.cuda.synchronize(qkv.device)
    return out


def dwconv_ln_swish(x: torch.Tensor, w_dw: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor,
                    plan: SemanticPlan, precision: str) -> torch.Tensor:
    lib = L.load()
    L.require_device(x.device)
    db = DeviceBatch(plan, x.device)
    out = torch.empty_like(x)
    L.check(lib.b2t_dwconv_ln_swish(x.data_ptr(), w_dw.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), db.byref(),
                                    out.data_ptr(), _prec(precision), L.stream_ptr()), 'dwconv_ln_swish')
    torch.cuda.synchronize(x.device)
    return out


def vq_argmin(x: torch.Tensor, codebook: torch.Tensor, apply_ln: bool = False,
              workspace: Optional[torch.Tensor] = None, impl: int = L.IMPL_AUTO, stats: Optional[dict] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [M, D] fp32, codebook [K, D] fp32 -> (int16 [M], int32 [M]).  `stats` (dict) receives the number
    of re-scanned rows and the largest observed fast-pass error (relative to the bound's scale)."""
    lib = L.load()
    L.require_device(x.device)
    M, D = x.shape
    K = codebook.shape[0]
    need = lib.b2t_vq_workspace_bytes(M, D, K)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    o16 = torch.empty(M, dtype=torch.int16, device=x.device)
    o32 = torch.empty(M, dtype=torch.int32, device=x.device)
    L.check(lib.b2t_vq_argmin(x.data_ptr(), x.stride(0), M, D, codebook.data_ptr(), None, K, int(apply_ln), impl,
                              o16.data_ptr(), o32.data_ptr(), workspace.data_ptr(), workspace.numel(),
                              L.stream_ptr()), 'vq_argmin')
    if stats is not None:
        nfb, err = C.c_uint(0), C.c_float(0)
        L.check(lib.b2t_vq_debug_stats(workspace.data_ptr(), M, D, K, C.byref(nfb), C.byref(err)), 'vq_debug_stats')
        stats['n_fallback'], stats['max_rel_err'] = nfb.value, err.value
    return o16, o32
