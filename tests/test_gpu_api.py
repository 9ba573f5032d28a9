"""GPU (-m gpu): the public AudioToken API (reference audiotoken/core.py signatures) end to end."""
import math
import os

import numpy as np
import pytest
import torch

from audiotoken_b200 import AudioToken, Tokenizers
from audiotoken_b200 import io as aio
from audiotoken_b200.weights import synthetic_waveform

pytestmark = pytest.mark.gpu
SR = 16000


def _pad_encode(tok, wave, chunk):
    """What the reference does for one segment: right-pad to chunk_size, mask, encoder(), trim."""
    n = wave.numel()
    L = chunk * SR
    x = torch.zeros(1, L)
    m = torch.zeros(1, L)
    x[0, :n] = wave
    m[0, :n] = 1
    out = tok.encoder(x.to(tok.device), m.to(tok.device)).cpu()
    return out[0, :, :math.ceil(n / SR * 50)].numpy()


def test_encode_batch_files_matches_per_segment_reference_semantics(cuda_device, tmp_path):
    chunk = 3
    lengths = {'a': SR * 3, 'b': SR * 7 + 1234, 'c.take2': 5000, 'd': SR * 3 + 3199, 'e': SR * 2 + 17}
    indir = tmp_path / 'in' / 'spk'
    indir.mkdir(parents=True)
    files = []
    for i, (name, n) in enumerate(lengths.items()):
        p = indir / f'{name}.wav'
        aio.write_wav(str(p), synthetic_waveform(70 + i, n, SR), SR)
        files.append(str(p))
    (indir / 'notes.txt').write_text('not audio')
    tok = AudioToken(tokenizer=Tokenizers.semantic_m, device='cuda:0', n_layers=2, synthetic_weights=True)
    out1 = tmp_path / 'out_files'
    tok.encode_batch_files(batch_size=2, outdir=str(out1), chunk_size=chunk, num_workers=2, audio_files=files)
    for f in files:
        stem = os.path.basename(f).split('.')[0]
        got = np.load(out1 / f'{stem}.npy')
        wave = aio.read_audio(f, SR)
        want = np.hstack([_pad_encode(tok, wave[0, s:s + chunk * SR], chunk)
                          for s in range(0, wave.shape[1], chunk * SR) if wave[0, s:s + chunk * SR].numel() >= 3200])
        assert got.dtype == np.int16 and got.shape == want.shape, (f, got.shape, want.shape)
        assert np.array_equal(got, want), f
    # directory mode keeps the relative layout; running twice is idempotent (no append on re-run)
    out2 = tmp_path / 'out_dir'
    for _ in range(2):
        tok.encode_batch_files(batch_size=3, outdir=str(out2), chunk_size=chunk, audio_dir=str(tmp_path / 'in'))
    for f in files:
        stem = os.path.splitext(os.path.basename(f))[0]
        a = np.load(out2 / 'spk' / f'{stem}.npy')
        b = np.load(out1 / (os.path.basename(f).split('.')[0] + '.npy'))
        assert np.array_equal(a, b)
    assert tok.last_stats['files'] == 5 and not tok.last_stats['errors']


def test_encode_single_inputs(cuda_device, tmp_path):
    tok = AudioToken(tokenizer='semantic_s', device='cuda:0', n_layers=1, synthetic_weights=True)
    assert tok.model_sample_rate == 16000 and tok.num_codebooks == 16
    x = synthetic_waveform(3, 16037, SR).unsqueeze(0)
    t = tok.encode(x)
    assert t.device.type == 'cpu' and t.dtype == torch.int16 and t.shape == (1, 1, 50)   # 49 rows + 1 pad row
    assert int(t.max()) < 1000                                                            # k-means codebook of 1000
    t2 = tok.encode(x.numpy())
    assert torch.equal(t, t2)
    p = tmp_path / 'x.wav'
    aio.write_wav(str(p), x[0], SR)
    t3 = tok.encode(p)
    assert t3.shape == t.shape
    with pytest.raises(ValueError):
        tok.encode(p, chunk_size=1)                           # trailing 37-sample chunk: no frame fits (the reference fails too)
    p2 = tmp_path / 'y.wav'
    aio.write_wav(str(p2), synthetic_waveform(4, 40000, SR), SR)
    t4 = tok.encode(p2, chunk_size=1)
    assert t4.dim() == 2 and t4.shape == (1, 50 + 50 + 24)    # [K, T_total] when chunked (reference core.py:175-179)
    with pytest.raises(AssertionError):
        tok.encode(torch.zeros(2, 16000))
    with pytest.raises(NotImplementedError):
        tok.encode(b'raw')
    with pytest.raises(ValueError):
        tok.encode(123)
    with pytest.raises(AssertionError):
        AudioToken('semantic_m', device='cuda:0', num_codebooks=3)


def test_file_loop_mixed_pcm_encodings(cuda_device, tmp_path):
    """One corpus with PCM16 and float32 payloads (decoded on the device) next to 32-bit and 8-bit PCM (decoded on the
    host, io.convert_chunks) and a stereo file (rejected, logged, skipped): the same audio gives the same tokens whatever
    its container encoding.  Samples are multiples of 256, so s / 32768 is exact in all four encodings."""
    from scipy.io import wavfile
    g = np.random.default_rng(5)
    n = SR * 4 + 777
    s16 = (np.round(8000 * np.sin(np.arange(n) * 0.05) + 1500 * g.standard_normal(n)) // 256 * 256).clip(-32768, 32512).astype(np.int16)
    indir = tmp_path / 'in'
    indir.mkdir()
    wavfile.write(indir / 'p16.wav', SR, s16)
    wavfile.write(indir / 'p32.wav', SR, s16.astype(np.int32) * 65536)
    wavfile.write(indir / 'p8.wav', SR, (s16.astype(np.int32) // 256 + 128).astype(np.uint8))
    wavfile.write(indir / 'f32.wav', SR, (s16.astype(np.float32) / 32768.0))
    wavfile.write(indir / 'stereo.wav', SR, np.stack([s16, s16], axis=1))
    tok = AudioToken(tokenizer=Tokenizers.semantic_m, device='cuda:0', n_layers=2, synthetic_weights=True)
    out = tmp_path / 'out'
    tok.encode_batch_files(batch_size=4, outdir=str(out), chunk_size=3, audio_dir=str(indir), num_workers=2)
    ref = np.load(out / 'p16.npy')
    assert ref.shape == (1, math.ceil(n / SR * 50))
    for name in ('p32', 'p8', 'f32'):
        assert np.array_equal(np.load(out / f'{name}.npy'), ref), name
    assert not (out / 'stereo.npy').exists()
    assert tok.last_stats['files'] == 4 and len(tok.last_stats['errors']) == 1
