"""GPU (-m gpu): device-side ingest (PCM decode + mono mix-down + sinc resampling, csrc/ingest.cu) through the C ABI
against the torchaudio goldens of the reference's convert_audio and against the numpy oracle."""
import math
import os

import numpy as np
import pytest
import torch

from audiotoken_b200 import io as aio
from audiotoken_b200 import ingest
from oracle import resample

pytestmark = pytest.mark.gpu


def test_resample_matches_reference_goldens(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, 'resample.npz'))
    k = 0
    while f'case{k}_meta' in g.files:
        sr, tgt, ch, n = (int(v) for v in g[f'case{k}_meta'])
        x = torch.from_numpy(g[f'case{k}_in'])
        y = ingest.convert_audio(x, sr, tgt, 'cuda:0').cpu().numpy()
        ref = g[f'case{k}_out']
        assert y.shape == ref.shape and y.dtype == np.float32
        assert float(np.abs(y - ref).max()) < 2e-5, (k, float(np.abs(y - ref).max()))
        if sr == tgt and ch == 1:
            assert np.array_equal(y, ref)
        k += 1
    assert k >= 7


def test_resample_pcm16_interleaved_stereo_and_edges(cuda_device):
    """int16 PCM in the interleaved [L, C] layout of a WAV file (passed as a transposed view), odd lengths, a 1-sample
    clip and an empty clip; against the numpy oracle on the dequantised samples."""
    g = torch.Generator().manual_seed(11)
    for sr, tgt, n in ((44100, 16000, 100003), (48000, 16000, 1), (16000, 24000, 777), (44100, 24000, 0)):
        pcm = torch.randint(-32768, 32767, (n, 2), generator=g, dtype=torch.int16)
        y = ingest.convert_audio(pcm.t(), sr, tgt, 'cuda:0').cpu().numpy()
        want = resample.convert_audio((pcm.t().float() / 32768.0).numpy(), sr, tgt) if n else np.zeros((1, 0), np.float32)
        assert y.shape == (1, math.ceil(n * tgt / sr)) == want.shape
        if n:
            assert float(np.abs(y - want).max()) < 2e-5
    with pytest.raises(RuntimeError):
        ingest.convert_audio(torch.zeros(3, 100), 44100, 16000, 'cuda:0')


def test_batch_reader_resamples_chunk_by_chunk_on_device(cuda_device, tmp_path):
    """reference utils.py:71-101: chunks are cut at the SOURCE rate and resampled one by one."""
    from scipy.io import wavfile
    sr, tgt, chunk = 22050, 16000, 1
    x = (np.random.default_rng(3).uniform(-0.5, 0.5, int(2.6 * sr)) * 32767).astype(np.int16)
    p = tmp_path / 'a.wav'
    wavfile.write(str(p), sr, x)
    chunks = aio.read_audio_chunks(str(p), tgt, chunk, device='cuda:0')
    host = aio.read_audio_chunks(str(p), tgt, chunk, device=None)
    assert len(chunks) == 3 and all(c.is_cuda for c in chunks)
    for i, (c, h) in enumerate(zip(chunks, host)):
        piece = x[i * sr:(i + 1) * sr].astype(np.float32)[None, :] / 32768.0
        want = resample.convert_audio(piece, sr, tgt)
        assert c.shape == want.shape == h.shape
        assert float(np.abs(c.cpu().numpy() - want).max()) < 2e-5
        assert float((c.cpu() - h).abs().max()) < 2e-5
