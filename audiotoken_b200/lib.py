"""ctypes binding of libb200tok.so (include/b200tok.h).

There is deliberately no fallback: if the library is missing or the device is not sm_100,
every entry point raises.  PyTorch is used only to own device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B2T_LIB_PATH') or os.path.join(_HERE, 'lib', 'libb200tok.so')   # override: developer A/B builds

PREC_BF16, PREC_FP32 = 0, 1
EPI_BIAS, EPI_BIAS_SWISH, EPI_RESID, EPI_GLU, EPI_BIAS_MASK, EPI_BIAS_GELU = 0, 1, 2, 3, 4, 5
IMPL_AUTO, IMPL_SIMT, IMPL_TENSOR, IMPL_MMA_SYNC = 0, 1, 2, 3

EXPORTS = [
    'b2t_version', 'b2t_last_error', 'b2t_device_check', 'b2t_set_option', 'b2t_fbank_logmel', 'b2t_fbank_stats',
    'b2t_fbank_stack_ln', 'b2t_layernorm', 'b2t_add_layernorm', 'b2t_gemm', 'b2t_relkey_attention', 'b2t_attention', 'b2t_dwconv_ln_swish',
    'b2t_vq_workspace_bytes', 'b2t_vq_argmin', 'b2t_vq_debug_stats', 'b2t_semantic_create', 'b2t_semantic_destroy',
    'b2t_semantic_set_tensor', 'b2t_semantic_workspace_bytes', 'b2t_semantic_encode',
    'b2t_last_launch_count', 'b2t_profile_enable', 'b2t_profile_read',
    'b2t_acoustic_create', 'b2t_acoustic_destroy', 'b2t_acoustic_set_tensor', 'b2t_acoustic_workspace_bytes',
    'b2t_acoustic_encode', 'b2t_rvq_encode', 'b2t_acoustic_profile_read', 'b2t_ingest_resample',
    'b2t_acoustic_decode_workspace_bytes', 'b2t_acoustic_decode', 'b2t_vq_ema_workspace_bytes', 'b2t_vq_ema_update',
    'b2t_hubert_create', 'b2t_hubert_destroy', 'b2t_hubert_set_tensor', 'b2t_hubert_workspace_bytes', 'b2t_hubert_encode',
    'b2t_debug_stage_sums', 'b2t_debug_stage_count', 'b2t_last_device_trap',
]


class B2TError(RuntimeError):
    pass


class Batch(C.Structure):
    """b2t_batch"""
    _fields_ = [('n_clips', C.c_int32), ('total_frames', C.c_int32), ('total_rows', C.c_int32),
                ('n_qtiles', C.c_int32), ('n_ctiles', C.c_int32), ('max_rows', C.c_int32),
                ('wave_off', C.c_void_p), ('frame_off', C.c_void_p), ('stack_frames', C.c_void_p),
                ('row_off', C.c_void_p), ('valid_rows', C.c_void_p), ('qtile_clip', C.c_void_p),
                ('qtile_q0', C.c_void_p), ('ctile_clip', C.c_void_p), ('ctile_t0', C.c_void_p),
                ('n_qtiles128', C.c_int32), ('qtile128_clip', C.c_void_p), ('qtile128_q0', C.c_void_p)]


class FbankTables(C.Structure):
    """b2t_fbank_tables"""
    _fields_ = [('window', C.c_void_p), ('mel_start', C.c_void_p), ('mel_count', C.c_void_p),
                ('mel_weight', C.c_void_p)]


class AcousticBatch(C.Structure):
    """b2t_acoustic_batch"""
    _fields_ = [('n_clips', C.c_int32), ('t_max', C.c_int32), ('total', C.c_int32 * 5), ('n_tiles', C.c_int32 * 5),
                ('wave_off', C.c_void_p), ('true_len', C.c_void_p), ('len', C.c_void_p * 5), ('off', C.c_void_p * 5),
                ('tile_clip', C.c_void_p * 5), ('tile_t0', C.c_void_p * 5), ('order', C.c_void_p),
                ('rank', C.c_void_p), ('toff', C.c_void_p), ('frames_host', C.c_void_p), ('aligned320', C.c_int32)]


class GemmArgs(C.Structure):
    """b2t_gemm_args"""
    _fields_ = [('A', C.c_void_p), ('lda', C.c_int32), ('W', C.c_void_p), ('bias', C.c_void_p),
                ('out', C.c_void_p), ('ldo', C.c_int32), ('resid', C.c_void_p), ('row_valid', C.c_void_p),
                ('M', C.c_int32), ('N', C.c_int32), ('K', C.c_int32), ('epilogue', C.c_int32),
                ('alpha', C.c_float), ('round_resid_bf16', C.c_int32), ('precision', C.c_int32),
                ('impl', C.c_int32)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2TError(f'{LIB_PATH} not found: build it with `python -m audiotoken_b200.build` '
                       '(there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.b2t_version.restype = i32
    lib.b2t_last_error.restype = C.c_char_p
    lib.b2t_device_check.argtypes = [i32]
    lib.b2t_set_option.argtypes = [C.c_char_p, i32]
    lib.b2t_fbank_logmel.argtypes = [vp, C.POINTER(Batch), C.POINTER(FbankTables), vp, i32, vp]
    lib.b2t_fbank_stats.argtypes = [vp, C.POINTER(Batch), vp, vp, vp]
    lib.b2t_fbank_stack_ln.argtypes = [vp, vp, vp, C.POINTER(Batch), vp, vp, vp, vp, vp, i32, vp]
    lib.b2t_layernorm.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.b2t_gemm.argtypes = [C.POINTER(GemmArgs), vp]
    lib.b2t_add_layernorm.argtypes = [vp, vp, C.c_float, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.b2t_relkey_attention.argtypes = [vp, vp, C.POINTER(Batch), vp, i32, i32, vp]
    lib.b2t_attention.argtypes = [vp, vp, C.POINTER(Batch), vp, i32, i32, i32, vp]
    lib.b2t_dwconv_ln_swish.argtypes = [vp, vp, vp, vp, C.POINTER(Batch), vp, i32, vp]
    lib.b2t_vq_workspace_bytes.argtypes = [i32, i32, i32]
    lib.b2t_vq_workspace_bytes.restype = sz
    lib.b2t_vq_argmin.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.b2t_vq_debug_stats.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_uint), C.POINTER(C.c_float)]
    lib.b2t_vq_ema_workspace_bytes.argtypes = [i32, i32, i32]
    lib.b2t_vq_ema_workspace_bytes.restype = sz
    f32 = C.c_float
    lib.b2t_vq_ema_update.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, f32, f32, f32, vp, vp, vp, sz, vp]
    lib.b2t_semantic_create.argtypes = [i32, i32, i32]
    lib.b2t_semantic_create.restype = vp
    lib.b2t_semantic_destroy.argtypes = [vp]
    lib.b2t_semantic_destroy.restype = None
    lib.b2t_semantic_set_tensor.argtypes = [vp, C.c_char_p, vp]
    lib.b2t_semantic_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.b2t_semantic_workspace_bytes.restype = sz
    lib.b2t_semantic_encode.argtypes = [vp, vp, C.POINTER(Batch), C.POINTER(FbankTables), vp, sz, vp, i32, vp, vp]
    lib.b2t_last_launch_count.restype = i32
    lib.b2t_last_device_trap.argtypes = [C.POINTER(C.c_uint32)]
    lib.b2t_last_device_trap.restype = i32
    lib.b2t_profile_enable.argtypes = [i32]
    lib.b2t_acoustic_create.restype = vp
    lib.b2t_acoustic_destroy.argtypes = [vp]
    lib.b2t_acoustic_destroy.restype = None
    lib.b2t_acoustic_set_tensor.argtypes = [vp, C.c_char_p, vp]
    lib.b2t_acoustic_workspace_bytes.argtypes = [C.POINTER(AcousticBatch), i32]
    lib.b2t_acoustic_workspace_bytes.restype = sz
    lib.b2t_acoustic_encode.argtypes = [vp, vp, C.POINTER(AcousticBatch), i32, i32, vp, sz, vp, vp, vp, vp]
    lib.b2t_acoustic_decode_workspace_bytes.argtypes = [C.POINTER(AcousticBatch)]
    lib.b2t_acoustic_decode_workspace_bytes.restype = sz
    lib.b2t_acoustic_decode.argtypes = [vp, vp, C.POINTER(AcousticBatch), i32, vp, sz, vp, vp, vp]
    ll = C.c_longlong
    lib.b2t_ingest_resample.argtypes = [vp, i32, ll, i32, ll, ll, vp, vp, vp, i32, i32, i32, i32, vp, ll, vp]
    lib.b2t_acoustic_profile_read.argtypes = [C.POINTER(C.c_float)]
    lib.b2t_rvq_encode.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.b2t_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_double)]
    lib.b2t_hubert_create.argtypes = [i32, i32, i32]
    lib.b2t_hubert_create.restype = vp
    lib.b2t_hubert_destroy.argtypes = [vp]
    lib.b2t_hubert_destroy.restype = None
    lib.b2t_hubert_set_tensor.argtypes = [vp, C.c_char_p, vp]
    lib.b2t_hubert_workspace_bytes.argtypes = [vp, vp]
    lib.b2t_hubert_workspace_bytes.restype = sz
    lib.b2t_hubert_encode.argtypes = [vp, vp, vp, vp, sz, i32, vp, i32, vp, vp, vp]
    lib.b2t_debug_stage_sums.argtypes = [vp, i32]
    _lib = lib
    return lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = load().b2t_last_error().decode('utf-8', 'replace')
        raise B2TError(f'{what} failed (status {rc}): {msg}')


def device_trap_text() -> str:
    """The record a register-critical kernel left in mapped host memory before it trapped (b2t_last_device_trap), or ''."""
    rec = (C.c_uint32 * 6)()
    if not load().b2t_last_device_trap(rec):
        return ''
    return (f'device trap record: site 0x{rec[0]:x} a={rec[1]} b={rec[2]} block=({rec[3]},{rec[4]}) thread={rec[5]}')


def require_device(device: torch.device) -> None:
    """Fail loudly unless `device` is a CUDA sm_100 device."""
    if device.type != 'cuda' or not torch.cuda.is_available():
        raise B2TError(f"device '{device}' is not a CUDA device: b200tok has no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(load().b2t_device_check(idx), 'b2t_device_check')


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def np_to_dev(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
