"""GPU (-m gpu): the reference's own semantic_s (mHuBERT-base + k-means, reference encoder.py:60-108) through the C ABI
against the HF-generated goldens (tests/golden/hubert.npz), the oracle (oracle/hubert.py) and HF HubertModel under CUDA
autocast.  Tolerances: fp32 within 1e-4 of the goldens / tokens >= 99.5 %; bf16 within 1e-2 of the autocast reference."""
import os

import numpy as np
import pytest
import torch

from audiotoken_b200.hubert import HubertEncoder, feat_lengths, plan_hubert
from audiotoken_b200.weights import synthetic_codebook, synthetic_hubert_state_dict, synthetic_waveform
from oracle import hubert as OH

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'hubert.npz'))
    lengths, total = [int(v) for v in g['lengths']], int(g['total'])
    raw = [synthetic_waveform(300 + i, n, 16000) for i, n in enumerate(lengths)]
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, (w, n) in enumerate(zip(raw, lengths)):
        wave[i, :n] = OH.processor_normalize(w)
        mask[i, :n] = 1
    return g, lengths, total, raw, wave, mask


def test_hubert_fp32_matches_hf_golden(cuda_device, golden_dir):
    """forward() on the padded batch: hidden states 0 / 1 / 11 of EVERY frame (padded frames included: zero after the
    projection, queries only), conv features of the valid frames, tokens."""
    g, lengths, total, raw, wave, mask = _golden(golden_dir)
    sd = synthetic_hubert_state_dict(0)
    cb = synthetic_codebook(1000, 768, seed=9)
    enc = HubertEncoder(device='cuda:0', precision='fp32', state_dict=sd, codebook=cb)
    toks, h11 = enc(wave.to(cuda_device), mask.to(cuda_device), tap_layer=11)
    torch.cuda.synchronize()
    assert toks.shape == (3, 1, 49) and toks.dtype == torch.int16
    assert rel(h11, g['h11']) < 1e-4
    for layer, key in ((0, 'h0'), (1, 'h1')):
        _, h = enc(wave.to(cuda_device), mask.to(cuda_device), tap_layer=layer)
        assert rel(h[:, ::3], g[key]) < 1e-5, layer
    agree = float((toks.cpu().numpy() == g['tokens']).mean())
    assert agree >= 0.995, agree
    # feature encoder output of the valid frames (GroupNorm statistics over the padded chunk)
    plan = plan_hubert(lengths, np.arange(3) * total, total)
    _, _, feats = enc.encode_plan(wave.to(cuda_device).view(-1), plan, want_feats=True)
    feats = feats.cpu().view(3, 49, 512)
    want = g['feats']                                        # [3, 49, 512][:, ::7, ::16]
    for i, tv in enumerate(feat_lengths(lengths)):
        rows = [r for r in range(0, 49, 7) if r < tv]
        got = feats[i, rows][:, ::16].numpy()
        assert np.abs(got - want[i, :len(rows)]).max() < 5e-5 * np.abs(want).max(), i


def test_hubert_packed_equals_padded_and_normalises_on_device(cuda_device, golden_dir):
    """encode_packed on RAW clips (device-side normalisation, ragged rows = the tokens the reference saves) gives the
    tokens of the padded batch; every clip alone gives the same tokens."""
    g, lengths, total, raw, wave, mask = _golden(golden_dir)
    sd = synthetic_hubert_state_dict(0)
    cb = synthetic_codebook(1000, 768, seed=9)
    for prec in ('fp32', 'bf16'):
        enc = HubertEncoder(device='cuda:0', precision=prec, state_dict=sd, codebook=cb)
        padded = enc(wave.to(cuda_device), mask.to(cuda_device)).cpu()
        rows = [min(-(-n // 320), 49) for n in lengths]                      # ceil(n / 320) tokens are saved
        packed = [t.cpu() for t in enc.encode_packed(raw, total, rows)]
        alone = [enc.encode_packed([c], total, [r])[0].cpu() for c, r in zip(raw, rows)]
        for i, r in enumerate(rows):
            assert packed[i].shape == (1, r)
            agree = float((packed[i][0] == padded[i, 0, :r]).float().mean())
            assert agree >= (1.0 if prec == 'fp32' else 0.97), (prec, i, agree)    # device fp64-sum normalisation vs torch fp32
            assert torch.equal(alone[i], packed[i]), (prec, i)


def test_hubert_bf16_matches_autocast_reference(cuda_device, golden_dir):
    """bf16 (tcgen05) against HF HubertModel under torch.amp.autocast on this GPU, and against the fp32 goldens."""
    from oracle import hf_reference as R
    g, lengths, total, raw, wave, mask = _golden(golden_dir)
    sd = synthetic_hubert_state_dict(0)
    cb = synthetic_codebook(1000, 768, seed=9)
    enc = HubertEncoder(device='cuda:0', precision='bf16', state_dict=sd, codebook=cb)
    toks, h11 = enc(wave.to(cuda_device), mask.to(cuda_device), tap_layer=11)
    torch.cuda.synchronize()
    ref_bf = R.hubert_reference(wave, mask, sd, 'cuda:0', autocast=True)
    ref_32 = R.hubert_reference(wave, mask, sd, 'cuda:0', autocast=False, tf32=False)
    assert rel(ref_32[11], g['h11']) < 1e-4                                   # the GPU fp32 run reproduces the CPU golden
    e_mine, e_ref, e_pair = rel(h11, ref_32[11]), rel(ref_bf[11], ref_32[11]), rel(h11, ref_bf[11])
    cbd = cb.double()

    def tok(h):
        e = torch.nn.functional.layer_norm(h.double(), (768,))
        return torch.cdist(e, cbd).argmin(-1)
    t32, tbf = tok(ref_32[11]), tok(ref_bf[11])
    mine = toks[:, 0].cpu().long()
    print(f'\nhubert bf16: hidden 11 vs fp32 reference: this library {e_mine:.4f}, HF autocast {e_ref:.4f}; this library vs HF autocast '
          f'{e_pair:.4f}; tokens: mine vs fp32 {float((mine == t32).float().mean()):.4f}, HF autocast vs fp32 '
          f'{float((tbf == t32).float().mean()):.4f}, mine vs HF autocast {float((mine == tbf).float().mean()):.4f}')
    assert e_pair < 1e-2, e_pair
    assert e_mine <= 1.2 * e_ref + 1e-3, (e_mine, e_ref)
    assert float((mine == t32).float().mean()) >= float((tbf == t32).float().mean()) - 0.02


def test_hubert_long_ragged_batch_runs_and_is_deterministic(cuda_device):
    """30 s / 10 s / odd lengths in one ragged batch (T up to 1499, 2 layers for speed), twice: identical tokens."""
    lens = [480000, 160000, 123457, 3200]
    clips = [synthetic_waveform(500 + i, n, 16000) for i, n in enumerate(lens)]
    enc = HubertEncoder(device='cuda:0', precision='bf16', n_layers=2)
    rows = [min(-(-n // 320), int(feat_lengths(480000))) for n in lens]
    a = [t.cpu() for t in enc.encode_packed(clips, 480000, rows)]
    b = [t.cpu() for t in enc.encode_packed(clips, 480000, rows)]
    for x, y, r in zip(a, b, rows):
        assert x.shape == (1, r) and torch.equal(x, y)
        assert int(x.min()) >= 0 and int(x.max()) < 1000


def test_audiotoken_semantic_s_hubert_api(cuda_device, tmp_path):
    """AudioToken('semantic_s', semantic_s_model='hubert'): encode(array), and encode_batch_files == per-chunk reference
    semantics (each streamed chunk normalised on its own, padded to chunk_size, ceil(n / 320) tokens saved)."""
    import math
    from audiotoken_b200 import AudioToken
    from audiotoken_b200 import io as aio
    sr, chunk = 16000, 2
    tok = AudioToken('semantic_s', device='cuda:0', semantic_s_model='hubert', synthetic_weights=True, n_layers=2, precision='fp32')
    x = synthetic_waveform(600, 20000, sr)
    single = tok.encode(x[None].numpy())
    assert single.shape == (1, 1, int(feat_lengths(20000))) and single.dtype == torch.int16
    files = []
    for i, n in enumerate((sr * 2, sr * 5 + 777, 4000)):
        p = tmp_path / f'h{i}.wav'
        aio.write_wav(str(p), synthetic_waveform(610 + i, n, sr), sr)
        files.append(str(p))
    out = tmp_path / 'out'
    tok.encode_batch_files(batch_size=2, outdir=str(out), chunk_size=chunk, audio_files=files, num_workers=2)
    assert tok.last_stats['files'] == 3 and not tok.last_stats['errors']
    enc = tok.encoder
    for f in files:
        got = np.load(out / (os.path.basename(f).split('.')[0] + '.npy'))
        wave = aio.read_audio(f, sr)
        want = []
        for a in range(0, wave.shape[1], chunk * sr):
            seg = wave[0, a:a + chunk * sr]
            if seg.numel() < 3200:
                continue
            # the reference: normalise the chunk, right-pad to chunk_size with a 0/1 mask, encoder(), keep ceil(n/320) tokens
            xb = torch.zeros(1, chunk * sr)
            mb = torch.zeros(1, chunk * sr)
            xb[0, :seg.numel()] = OH.processor_normalize(seg)
            mb[0, :seg.numel()] = 1
            t = enc(xb.to(cuda_device), mb.to(cuda_device)).cpu()
            want.append(t[0, :, :math.ceil(seg.numel() / sr * 50)].numpy())
        want = np.hstack(want)
        assert got.shape == want.shape and got.dtype == np.int16, (f, got.shape, want.shape)
        assert (got == want).mean() >= 0.99, f          # device fp64-sum normalisation vs torch fp32 before the argmin
