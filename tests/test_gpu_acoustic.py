"""GPU (-m gpu): acoustic path (SEANet encoder + LSTM + RVQ) through the C ABI vs the EnCodec stand-in goldens
and the oracle.  fp32 numerics: embeddings within 1e-4 relative, codes >= 99.5 % (RVQ stages compound, so a
single near-tie flips the later stages of that frame)."""
import math
import os

import numpy as np
import pytest
import torch

from audiotoken_b200 import AudioToken, Tokenizers
from audiotoken_b200 import io as aio
from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
from audiotoken_b200.weights import synthetic_encodec_state_dict, synthetic_waveform
from oracle import seanet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def enc(cuda_device):
    return AcousticEncoder(device='cuda:0', state_dict=synthetic_encodec_state_dict(0), precision='fp32')


@pytest.fixture(scope='module')
def enc16(cuda_device):
    return AcousticEncoder(device='cuda:0', state_dict=synthetic_encodec_state_dict(0), precision='bf16')


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_acoustic_matches_golden(enc, cuda_device, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, 'acoustic.npz'))
    lengths = g[f'lengths_{tag}']
    w = torch.stack([synthetic_waveform(20 + i, int(n), 24000) for i, n in enumerate(lengths)])
    codes, emb = enc(w.to(cuda_device), None, want_emb=True)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g[f'emb_{tag}'])
    assert emb.shape == ref.shape and codes.dtype == torch.int16
    err = float((emb.cpu() - ref).norm() / ref.norm())
    assert err < 1e-4, err
    agree = float((codes.cpu().numpy() == g[f'codes_{tag}']).mean())
    assert agree >= 0.995, agree
    # given the kernel's own embeddings, the RVQ stage is exact (fp64 oracle on the same fp32 residuals)
    want = seanet.rvq_codes(emb.cpu().float(), synthetic_encodec_state_dict(0), 16).transpose(0, 1)
    assert torch.equal(codes.cpu().long(), want)


@pytest.mark.parametrize('tag', ['a', 'c'])
def test_acoustic_bf16_tensor_path_matches_golden(enc16, cuda_device, golden_dir, tag):
    """tcgen05 encoder (bf16 operands, fp32 accumulation = the reference's GPU autocast numerics) against the fp32
    EnCodec stand-in goldens: embeddings within 1e-2 relative (north star, bf16); the RVQ stage is exact given the
    kernel's own embeddings; the first codebooks agree with the fp32 reference up to the embedding noise."""
    g = np.load(os.path.join(golden_dir, 'acoustic.npz'))
    lengths = g[f'lengths_{tag}']
    w = torch.stack([synthetic_waveform(20 + i, int(n), 24000) for i, n in enumerate(lengths)])
    codes, emb = enc16(w.to(cuda_device), None, want_emb=True)
    torch.cuda.synchronize()
    assert enc16.last_precision == 'bf16'
    ref = torch.from_numpy(g[f'emb_{tag}'])
    assert emb.shape == ref.shape and codes.dtype == torch.int16
    err = float((emb.cpu() - ref).norm() / ref.norm())
    print(f'bf16 tensor path, golden {tag}: embedding rel err {err:.3e}')
    assert err < 1e-2, err
    want = seanet.rvq_codes(emb.cpu().float(), synthetic_encodec_state_dict(0), 16).transpose(0, 1)
    assert torch.equal(codes.cpu().long(), want)
    agree0 = float((codes.cpu().numpy()[:, 0] == g[f'codes_{tag}'][:, 0]).mean())
    print(f'first-codebook agreement with the fp32 reference: {agree0:.4f}')
    assert agree0 >= 0.9, agree0


def test_acoustic_bf16_matches_fp32_kernels_incl_tiny_clips(enc, enc16, cuda_device):
    """Ragged batch with 1-frame and 6-frame clips (halo rows the mirror cannot fill must read as zeros: the
    short-input rule of EncodecConv1d._pad1d) — the tensor path against the fp32 CUDA-core path."""
    lengths = [320, 320 * 6, 320 * 7, 320 * 40, 320 * 2, 320 * 133]
    clips = [synthetic_waveform(80 + i, n, 24000) for i, n in enumerate(lengths)]
    offs = np.zeros(len(clips), dtype=np.int64)
    offs[1:] = np.cumsum(lengths)[:-1]
    wave = torch.cat(clips).to(cuda_device)
    plan = plan_acoustic(lengths, offs, lengths)
    _, e32 = enc.encode_plan(wave, plan, want_emb=True)
    _, e16 = enc16.encode_plan(wave, plan, want_emb=True)
    torch.cuda.synchronize()
    assert enc16.last_precision == 'bf16' and enc.last_precision == 'fp32'
    fo = plan.offs[4]
    for i in range(len(lengths)):
        a, b = e32[fo[i]:fo[i + 1]].cpu(), e16[fo[i]:fo[i + 1]].cpu()
        err = float((a - b).norm() / a.norm())
        assert err < 1.5e-2, (i, err)


def test_rvq_tensor_exact(cuda_device):
    """Residual-VQ stage alone: tcgen05 kernel == CUDA-core kernel == fp64 oracle, on random residuals, on rows that
    ARE codewords (zero residual afterwards), with duplicated codewords (first index wins) and a ragged row count."""
    from audiotoken_b200 import lib as L
    sd = synthetic_encodec_state_dict(0)
    sd = {k: v.clone() for k, v in sd.items()}
    e0 = sd['quantizer.layers.0.codebook.embed']
    e0[700] = e0[13]                         # exact duplicate: index 13 must win
    e0[901] = e0[900] * (1 + 2 ** -20)       # near-duplicate inside the error band
    enc = AcousticEncoder(device='cuda:0', state_dict=sd, precision='bf16')
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(1000, 128, generator=g) * 0.9
    emb[:64] = e0[torch.arange(64) * 3 % 1024]                 # exact codewords
    emb[64:96] = e0[13] + 1e-4 * torch.randn(32, 128, generator=g)
    emb[96:128] = e0[900] + 1e-5 * torch.randn(32, 128, generator=g)
    want = seanet.rvq_codes(emb.t().unsqueeze(0), sd, 16)[:, 0]            # [16, 1000]
    dev = emb.to(cuda_device).contiguous()
    enc.rvq_stats()
    tc = enc.rvq_encode(dev, L.IMPL_TENSOR).cpu().long()
    rescored, rescans = enc.rvq_stats()
    simt = enc.rvq_encode(dev, L.IMPL_SIMT).cpu().long()
    print(f'tensor RVQ: {rescored} fp64 re-scores, {rescans} exhaustive re-scans of {16 * 1000} (row, stage) pairs')
    assert torch.equal(simt, want)
    assert torch.equal(tc, want)
    assert rescored > 0


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_acoustic_packed_equals_padded(enc, enc16, cuda_device, prec):
    enc = enc if prec == 'fp32' else enc16
    lengths = [24000, 7777, 3200, 15001]
    clips = [synthetic_waveform(40 + i, n, 24000) for i, n in enumerate(lengths)]
    total = 24000
    wave = torch.zeros(len(lengths), total)
    for i, c in enumerate(clips):
        wave[i, :c.numel()] = c
    padded = enc(wave.to(cuda_device), None).cpu()
    rows = [math.ceil(n / 24000 * 75) for n in lengths]
    packed = [t.cpu() for t in enc.encode_packed(clips, total, rows)]
    for i, r in enumerate(rows):
        assert packed[i].shape == (16, r)
        assert torch.equal(packed[i], padded[i, :, :r]), i
    alone = enc.encode_packed([clips[1]], total, [rows[1]])[0].cpu()
    assert torch.equal(alone, packed[1])


def test_acoustic_audiotoken_api(cuda_device, tmp_path):
    tok = AudioToken(tokenizer=Tokenizers.acoustic, device='cuda:0', num_codebooks=8, synthetic_weights=True)
    assert tok.model_sample_rate == 24000
    x = synthetic_waveform(5, 24000, 24000).unsqueeze(0)
    t = tok.encode(x)
    assert t.shape == (1, 8, 75) and t.dtype == torch.int16 and t.device.type == 'cpu'
    full = AudioToken(tokenizer='acoustic', device='cuda:0', synthetic_weights=True).encode(x)
    assert full.shape == (1, 16, 75) and torch.equal(full[:, :8], t)       # RVQ prefix property
    files = []
    for i, n in enumerate((24000 * 2 + 4000, 9000, 24000 + 100)):
        p = tmp_path / f'f{i}.wav'
        aio.write_wav(str(p), synthetic_waveform(60 + i, n, 24000), 24000)
        files.append(str(p))
    tok.encode_batch_files(batch_size=2, outdir=str(tmp_path / 'o'), chunk_size=1, audio_files=files)
    a = np.load(tmp_path / 'o' / 'f0.npy')
    assert a.dtype == np.int16 and a.shape == (8, 75 + 75 + 13)
    b = np.load(tmp_path / 'o' / 'f1.npy')
    assert b.shape == (8, math.ceil(9000 / 24000 * 75))
    # a trailing 100-sample chunk is below the 3200-sample minimum and is skipped (reference datasets.py:95-97)
    assert np.load(tmp_path / 'o' / 'f2.npy').shape == (8, 75)


# ---------------------------------------------------------------------------------------- decode half
@pytest.mark.parametrize('tag', ['a', 'c'])
def test_acoustic_decode_matches_golden(cuda_device, golden_dir, tag):
    """reference decoder.py:62-76 (quantizer.decode + SEANet decoder) through b2t_acoustic_decode against the
    EnCodec stand-in goldens: waveform within 1e-4 relative (fp32)."""
    from audiotoken_b200.acoustic import AcousticDecoder
    g = np.load(os.path.join(golden_dir, 'acoustic.npz'))
    dec = AcousticDecoder(device='cuda:0', state_dict=synthetic_encodec_state_dict(0))
    codes = torch.from_numpy(g[f'codes_{tag}'])                            # [B, 16, T]
    wav = dec(codes.to(cuda_device))
    torch.cuda.synchronize()
    ref = torch.from_numpy(g[f'dec_{tag}']).reshape(1, -1)
    assert wav.shape == ref.shape and wav.dtype == torch.float32
    err = float((wav.cpu() - ref).norm() / ref.norm())
    assert err < 1e-4, err
    # ragged: the same clips decoded as one packed batch of different lengths == each clip alone (causal decoder)
    B, K, T = codes.shape
    if B == 2:
        c0, c1 = codes[0, :, :40], codes[1]
        packed = torch.cat([c0, c1], dim=1).to(torch.int16).contiguous().to(cuda_device)
        w = dec.decode_packed(packed, [40, T]).cpu()
        alone0 = dec(c0.unsqueeze(0).to(cuda_device)).cpu().reshape(-1)
        alone1 = dec(c1.unsqueeze(0).to(cuda_device)).cpu().reshape(-1)
        assert torch.equal(w[:40 * 320], alone0) and torch.equal(w[40 * 320:], alone1)
        # causality: the first 40 frames of a clip decode to the first 12800 samples of the full decode
        assert torch.equal(alone0, ref.reshape(B, -1)[0, :40 * 320]) or float((alone0 - ref.reshape(B, -1)[0, :40 * 320]).abs().max()) < 1e-4


def test_acoustic_audiotoken_encode_decode_roundtrip_api(cuda_device):
    """AudioToken.decode for the acoustic tokenizer (reference core.py:317-357): shapes, dtype, determinism."""
    tok = AudioToken(tokenizer=Tokenizers.acoustic, device='cuda:0', num_codebooks=8, synthetic_weights=True)
    x = synthetic_waveform(5, 24000, 24000).unsqueeze(0)
    t = tok.encode(x)                                                     # [1, 8, 75]
    y = tok.decode(t)
    assert y.shape == (1, 24000) and y.dtype == torch.float32 and y.device.type == 'cpu'
    assert torch.equal(y, tok.decode(t.numpy()))
    with pytest.raises(NotImplementedError):
        AudioToken(tokenizer=Tokenizers.semantic_m, device='cuda:0', n_layers=1, synthetic_weights=True).decode(torch.zeros(1, 1, 10, dtype=torch.long))
