"""b200tok: B200-native drop-in for the encode path of cmeraki/audiotoken.

Re-exports the reference's public names (audiotoken/__init__.py:1-3).  Unlike the reference nothing
global is switched at import time (its TF32 / matmul-precision flags, __init__.py:6-9, only matter for
PyTorch library GEMMs, which this package does not use).
"""
from .configs import AUDIO_EXTS, TAR_EXTS, ZIP_EXTS, Tokenizers
from .core import AudioToken
from .io import read_audio

__all__ = ['AudioToken', 'Tokenizers', 'read_audio', 'AUDIO_EXTS', 'TAR_EXTS', 'ZIP_EXTS']
__version__ = '0.1.0'
