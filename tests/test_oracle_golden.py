"""CPU: the oracle (oracle/) against fixtures generated from the reference itself
(tests/golden/make_golden.py ran /root/reference's processors.py + attention patch on HF models)."""
import os

import numpy as np
import torch

from audiotoken_b200.weights import synthetic_codebook, synthetic_w2vbert_state_dict, synthetic_waveform
from oracle import conformer, fbank, quantize


def _clips(lengths, total):
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        wave[i, :n] = synthetic_waveform(i, int(n), 16000)
        mask[i, :n] = 1
    return wave, mask


def test_mel_filters_and_window_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fbank.npz'))
    assert np.array_equal(fbank.mel_filters().numpy(), g['mel_filters'])
    assert np.array_equal(fbank.povey_window().numpy(), g['window'])
    m = g['mel_filters']
    assert m.shape == (257, 80) and not m[256].any() and not m[0].any()
    assert ((m[:256] != 0).sum(axis=1) <= 2).all()          # each FFT bin feeds at most two filters


def test_features_match_reference_ragged_batch(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fbank.npz'))
    wave, mask = _clips(g['lengths'], int(g['total']))
    x, am = fbank.features(wave, mask)
    np.testing.assert_allclose(x.numpy(), g['input_features'], rtol=0, atol=1e-5)
    assert np.array_equal(am.numpy(), g['attention_mask'])
    assert am.sum(1).tolist() == [49.0, 34.0, 24.0, 9.0]


def test_features_match_reference_single_clips(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fbank_single.npz'))
    for n in (4800, 16037):
        wave, mask = _clips([n], n)
        x, am = fbank.features(wave, mask)
        np.testing.assert_allclose(x.numpy(), g[f'feat_{n}'], rtol=0, atol=1e-5)
        assert np.array_equal(am.numpy(), g[f'mask_{n}'])


def _run_conformer(golden_dir, tag, n_layers):
    g = np.load(os.path.join(golden_dir, 'fbank.npz'))
    c = np.load(os.path.join(golden_dir, f'conformer_{tag}.npz'))
    sd = synthetic_w2vbert_state_dict(n_layers, seed=0)
    hs = conformer.hidden_states(torch.from_numpy(g['input_features']), torch.from_numpy(g['attention_mask']),
                                 sd, n_layers)
    return hs, c


def test_conformer_2_layers_match_reference(golden_dir):
    hs, c = _run_conformer(golden_dir, 'l2', 2)
    for i, key in enumerate(('hidden_0', 'hidden_1', 'hidden_last')):
        ref = c[key]
        err = np.linalg.norm(hs[i].numpy() - ref) / np.linalg.norm(ref)
        assert err < 1e-5, (key, err)
    emb = conformer.final_embedding(hs[2]).reshape(-1, 1024)
    idx, tie = quantize.nearest_centroid(emb, synthetic_codebook(2048, 1024, 4))
    assert not tie.any()
    assert np.array_equal(idx.numpy().reshape(c['tokens'].shape), c['tokens'])


def test_conformer_19_layers_match_reference(golden_dir):
    hs, c = _run_conformer(golden_dir, 'l19', 19)
    ref = c['hidden_last']
    err = np.linalg.norm(hs[19].numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-4, err
    emb = conformer.final_embedding(hs[19]).reshape(-1, 1024)
    idx, _ = quantize.nearest_centroid(emb, synthetic_codebook(2048, 1024, 4))
    agree = (idx.numpy().reshape(c['tokens'].shape) == c['tokens']).mean()
    assert agree >= 0.995, agree


def test_nearest_centroid_matches_reference_expression():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(512, 256, generator=g)
    cb = torch.randn(1000, 256, generator=g)
    idx, tie = quantize.nearest_centroid(x, cb)
    ref = quantize.kmeans_assign_fp32(x, cb)                 # encoder.py:100-101 verbatim
    assert (idx[~tie] == ref[~tie]).float().mean() > 0.999
    cb2 = torch.cat([cb, cb[:5]], 0)                         # duplicates: first index wins
    idx2, _ = quantize.nearest_centroid(x, cb2)
    assert (idx2 < 1000).all()


def test_rvq_residual_property():
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(300, 128, generator=g)
    cbs = torch.randn(16, 1024, 128, generator=g)
    codes = quantize.rvq_encode(emb, cbs, 16)
    assert codes.shape == (16, 300)
    r = emb.double().clone()
    prev = (r * r).sum(1)
    for q in range(16):
        r = r - cbs[q].double()[codes[q]]
        cur = (r * r).sum(1)
        # the chosen code is the nearest one, so no other single code gives a smaller residual
        alt = ((r + cbs[q].double()[codes[q]]).unsqueeze(1) - cbs[q].double().unsqueeze(0)).pow(2).sum(-1).min(1).values
        assert torch.allclose(cur, alt)
        prev = cur
    assert quantize.num_quantizers(12.0) == 16 and quantize.num_quantizers(1.5) == 2


def test_seanet_oracle_matches_encodec_standin(golden_dir):
    """oracle/seanet.py against HF EncodecModel outputs (embeddings + RVQ-16 codes), incl. an odd length
    (reflect extra padding on the right) and a 333-sample clip (short-input rule of the reflect pad)."""
    from audiotoken_b200.weights import synthetic_encodec_state_dict
    from oracle import seanet
    g = np.load(os.path.join(golden_dir, 'acoustic.npz'))
    sd = synthetic_encodec_state_dict(0)
    for tag in 'abcd':
        lengths = g[f'lengths_{tag}']
        w = torch.stack([synthetic_waveform(20 + i, int(n), 24000) for i, n in enumerate(lengths)])
        emb = seanet.encoder(w, sd)
        ref = torch.from_numpy(g[f'emb_{tag}'])
        assert emb.shape == ref.shape
        assert float((emb - ref).abs().max()) < 2e-5
        codes = seanet.rvq_codes(emb, sd, 16).transpose(0, 1)
        agree = float((codes.numpy() == g[f'codes_{tag}']).mean())
        assert agree >= 0.995, (tag, agree)
        ref32 = seanet.rvq_codes_reference_fp32(emb, sd, 16).transpose(0, 1)
        assert float((ref32 == codes).float().mean()) >= 0.995


# ------------------------------------------------------------------------------ ingest (resampling)
def test_resample_oracle_matches_torchaudio_goldens(golden_dir):
    """oracle/resample.py against outputs of the reference's convert_audio (torchaudio.transforms.Resample defaults):
    mono / stereo, down- and up-sampling, equal rates."""
    from oracle import resample
    g = np.load(os.path.join(golden_dir, 'resample.npz'))
    k = 0
    while f'case{k}_meta' in g.files:
        sr, tgt, ch, n = (int(v) for v in g[f'case{k}_meta'])
        y = resample.convert_audio(g[f'case{k}_in'], sr, tgt)
        ref = g[f'case{k}_out']
        assert y.shape == ref.shape, (k, y.shape, ref.shape)
        assert float(np.abs(y - ref).max()) < 2e-5, k          # fp32 summation order of a <= 475-tap filter
        k += 1
    assert k >= 7
    # published structure of the filter bank: gcd-reduced rates, 2*width + orig taps per phase
    kern, width = resample.sinc_kernel(44100, 16000)
    assert kern.shape == (160, 2 * width + 441) and width == 17


def test_acoustic_decoder_oracle_matches_encodec_golden(golden_dir):
    """oracle/seanet.py decode half (codeword sum + SEANet decoder with causal transposed convs) against HF
    EncodecModel.quantizer.decode / .decoder on the synthetic weights (reference decoder.py:67-68)."""
    from audiotoken_b200.weights import synthetic_encodec_state_dict
    from oracle import seanet
    g = np.load(os.path.join(golden_dir, 'acoustic.npz'))
    sd = synthetic_encodec_state_dict(0)
    for tag in ('a', 'c'):
        codes = torch.from_numpy(g[f'codes_{tag}']).long().transpose(0, 1)      # [16, B, T]
        deq = seanet.rvq_decode(codes, sd)
        assert float((deq - torch.from_numpy(g[f'deq_{tag}'])).abs().max()) < 1e-6
        wav = seanet.decoder(deq, sd)
        ref = torch.from_numpy(g[f'dec_{tag}'])
        assert wav.shape == ref.shape
        assert float((wav - ref).norm() / ref.norm()) < 1e-5


def test_vq_ema_oracle_equals_onehot_formulation():
    """oracle/quantize.py::vq_ema_train_step against the one-hot / matmul formulation the third-party class
    publishes (one_hot -> sum, x^T onehot, mul_(decay).add_(new*(1-decay)), laplace smoothing), written out
    independently here; plus two properties: decay = 0 gives the (smoothed) cluster means, empty clusters shrink."""
    import torch
    from oracle import quantize
    g = torch.Generator().manual_seed(11)
    M, D, K = 600, 32, 50
    x = torch.randn(M, D, generator=g, dtype=torch.float64)
    embed = torch.randn(K, D, generator=g, dtype=torch.float64)
    embed[7] = 100.0                                   # never selected
    avg0, cs0 = embed.clone(), torch.full((K,), 3.0, dtype=torch.float64)
    idx, loss, e, avg, cs = quantize.vq_ema_train_step(x, embed, avg0, cs0, decay=0.8, eps=1e-5)
    dist = -torch.cdist(x, embed)
    ind = dist.argmax(-1)
    onehot = torch.nn.functional.one_hot(ind, K).to(torch.float64)
    cs_ref = cs0.clone().mul_(0.8).add_(onehot.sum(0) * 0.2)
    avg_ref = avg0.clone().mul_(0.8).add_((x.t() @ onehot).t() * 0.2)
    sm = (cs_ref + 1e-5) / (cs_ref.sum() + K * 1e-5) * cs_ref.sum()
    assert torch.equal(idx, ind)
    assert torch.allclose(cs, cs_ref, rtol=1e-13) and torch.allclose(avg, avg_ref, rtol=1e-12, atol=1e-13)
    assert torch.allclose(e, avg_ref / sm.unsqueeze(1), rtol=1e-12, atol=1e-13)
    assert abs(loss - float(((embed[ind] - x) ** 2).mean())) < 1e-12
    assert cs[7] == 3.0 * 0.8
    _, _, e0, _, cs00 = quantize.vq_ema_train_step(x, embed, avg0, cs0, decay=0.0, eps=1e-5)
    used = torch.unique(ind)
    means = torch.stack([x[ind == k].mean(0) for k in used])
    assert torch.allclose(e0[used], means, rtol=1e-4)


def test_hubert_oracle_matches_hf_golden(golden_dir):
    """oracle/hubert.py (the reference's own semantic_s = mHuBERT-base + k-means) against outputs of
    HF HubertModel / Wav2Vec2FeatureExtractor called the way reference encoder.py:88-103 calls them
    (tests/golden/make_golden_hubert.py): normalisation, conv features (GroupNorm over the padded chunk), hidden
    states 0 / 1 / 11 / 12 with padding masks, tokens."""
    import numpy as np
    import torch
    from audiotoken_b200.weights import synthetic_codebook, synthetic_hubert_state_dict, synthetic_waveform
    from oracle import hubert
    g = np.load(os.path.join(golden_dir, 'hubert.npz'))
    lengths, total = [int(v) for v in g['lengths']], int(g['total'])
    sd = synthetic_hubert_state_dict(0)
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        w = hubert.processor_normalize(synthetic_waveform(300 + i, n, 16000))
        if i == 0:
            assert np.abs(w[:64].numpy() - g['norm0']).max() < 2e-6
        if i == 1:
            assert abs(float(w.mean()) - g['norm1_stats'][0]) < 1e-6 and abs(float(w.std(unbiased=False)) - g['norm1_stats'][1]) < 1e-5
        wave[i, :n] = w
        mask[i, :n] = 1
    assert [int(v) for v in hubert.feat_lengths(torch.tensor(lengths))] == [49, 34, 12]
    feats = hubert.feature_encoder(wave, sd)
    assert np.abs(feats[:, ::7, ::16].numpy() - g['feats']).max() < 2e-5 * np.abs(g['feats']).max()
    hs, fm = hubert.hidden_states(wave, mask, sd)
    assert fm.sum(1).tolist() == [49, 34, 12]

    def rel(a, b):
        a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
        return float((a - b).norm() / b.norm())
    assert rel(hs[0][:, ::3], g['h0']) < 2e-6
    assert rel(hs[1][:, ::3], g['h1']) < 5e-6
    assert rel(hs[11], g['h11']) < 2e-5
    assert rel(hs[12][:, ::3], g['h12']) < 2e-5
    tok = hubert.tokens(hs, synthetic_codebook(1000, 768, seed=9))
    want = torch.as_tensor(g['tokens'])
    assert tok.shape == want.shape == (3, 1, 49) and tok.dtype == torch.int16
    valid = fm[:, None, :]
    assert float((tok[valid] == want[valid]).float().mean()) >= 0.995


def test_hubert_ragged_equals_padded(golden_dir):
    """The packed-batch restatement (no padding materialised; GroupNorm statistics over the padded frame count) equals
    the padded computation on every valid frame, and a different padded length changes the result (the mHuBERT front
    end, unlike the w2v-BERT fbank, is NOT padding-invariant)."""
    import numpy as np
    import torch
    from audiotoken_b200.weights import synthetic_hubert_state_dict, synthetic_waveform
    from oracle import hubert
    g = np.load(os.path.join(golden_dir, 'hubert.npz'))
    lengths, total = [int(v) for v in g['lengths']], int(g['total'])
    sd = synthetic_hubert_state_dict(0)
    clips = [hubert.processor_normalize(synthetic_waveform(300 + i, n, 16000)) for i, n in enumerate(lengths)]
    rag = hubert.hidden_states_ragged(clips, total, sd)
    for i, hs in enumerate(rag):
        tv = hs[11].shape[0]
        ref = torch.as_tensor(g['h11'])[i, :tv].double()
        assert float((hs[11].double() - ref).norm() / ref.norm()) < 2e-5, i
    other = hubert.hidden_states_ragged(clips[1:2], 2 * total, sd)[0]
    ref = torch.as_tensor(g['h11'])[1, :other[11].shape[0]].double()
    assert float((other[11].double() - ref).norm() / ref.norm()) > 1e-3


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _long_batch(golden_dir):
    g = np.load(os.path.join(golden_dir, 'conformer_long_l2.npz'))
    lengths = [int(v) for v in g['lengths']]
    total = int(g['total'])
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        wave[i, :n] = synthetic_waveform(i, n, 16000)
        mask[i, :n] = 1
    return g, wave, mask


def test_hf_reference_matches_long_golden(golden_dir):
    """The secondary oracle (HF model + the restated relative-key SDPA attention + the device-generic front end) on the
    CPU in fp32 against the fixture the REAL reference files produced at BASELINE shapes (10 s / 30 s clips padded to
    30 s, T = 1500): pins oracle/hf_reference.py before the GPU tests run it under CUDA autocast."""
    from oracle import hf_reference as R
    g, wave, mask = _long_batch(golden_dir)
    emb, am, hid = R.reference_embeddings(wave, mask, synthetic_w2vbert_state_dict(2, 0), 2, 'cpu', autocast=False)
    assert np.array_equal(am.numpy().astype(np.uint8), g['attention_mask'])
    for i in range(3):
        rows = g[f'rows_{i}']
        assert rel_err(hid[i, rows], torch.from_numpy(g[f'hidden_{i}'])) < 1e-6
    tok = R.vq_eval_tokens(emb, synthetic_codebook(2048, 1024, 4)).view(3, -1).numpy()
    m = g['attention_mask'].astype(bool)
    assert (tok[m] == g['tokens'][m]).mean() >= 0.999


def test_conformer_oracle_matches_long_golden(golden_dir):
    """The plain restatement (oracle/conformer.py) at T = 1500: the -64 / +8 clamp of the relative-key bias and key
    masking over >1000 padded keys against the real reference's output."""
    g, wave, mask = _long_batch(golden_dir)
    sd = synthetic_w2vbert_state_dict(2, 0)
    feats, am = fbank.features(wave, mask)
    for i in range(3):
        rows = g[f'rows_{i}']
        assert float((feats[i, rows] - torch.from_numpy(g[f'features_{i}'])).abs().max()) == 0.0
    hs = conformer.hidden_states(feats, am, sd, 2)
    for i in range(3):
        rows = g[f'rows_{i}']
        assert rel_err(hs[2][i, rows], torch.from_numpy(g[f'hidden_{i}'])) < 1e-5
    idx, _ = quantize.nearest_centroid(conformer.final_embedding(hs[2]).reshape(-1, 1024), synthetic_codebook(2048, 1024, 4))
    m = g['attention_mask'].astype(bool)
    assert (idx.view(3, -1).numpy()[m] == g['tokens'][m]).mean() >= 0.999
