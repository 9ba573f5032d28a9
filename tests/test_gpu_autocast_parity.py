"""GPU (-m gpu): the timed mode (bf16, 19 layers) against THE REFERENCE AS DEPLOYED, run on the same B200:
HF ``Wav2Vec2BertModel`` + the reference's relative-key SDPA attention + its processor under
``torch.amp.autocast('cuda', bfloat16)`` with TF32 allowed (reference encoder.py:163-184, __init__.py:6-9) —
oracle/hf_reference.py, pinned on the CPU against fixtures made by the real reference files.

North star (BASELINE.json): token agreement >= 99.5 % against the reference's own PyTorch path, embeddings within 1e-2
in bf16.  Random-init weights amplify bf16 rounding through 19 layers until two CORRECT bf16 implementations disagree on
more than 0.5 % of tokens, so the figures are reported the way SURVEY 7.3 (c/d) prescribes:

  (i)   this library (bf16) vs reference-bf16 (autocast)                 token agreement
  (ii)  reference-bf16 vs reference-fp32                                  = the noise floor of the reference itself
  (i')  this library (bf16) vs reference-fp32                             (compare with (ii): same truth, same precision)
  (iii) agreement on the rows whose fp64 top-2 margin exceeds the measured bf16 embedding noise of the reference

Asserted: (iii) >= 99.5 %, (i') >= (ii) - 0.5 pt, (i) >= (ii) - 1.0 pt, embedding error of this library vs
reference-fp32 <= 1.15 x that of reference-bf16.
"""
import os

import numpy as np
import pytest
import torch

from audiotoken_b200.encoder import Wav2VecBertEncoder
from audiotoken_b200.weights import (data_derived_codebook, synthetic_encodec_state_dict, synthetic_w2vbert_state_dict,
                                     synthetic_waveform)
from oracle import hf_reference as R
from oracle import quantize

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def _clips(lengths, total, sr=16000, first=0):
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        wave[i, :n] = synthetic_waveform(first + i, n, sr)
        mask[i, :n] = 1
    return wave, mask


def _margins(emb, codebook):
    """fp64 top-2 of every row: (best index, distance of the row to the bisector plane of its two nearest codewords)."""
    x = emb.double()
    c = codebook.double()
    d2 = (x * x).sum(1, keepdim=True) - 2.0 * (x @ c.t()) + (c * c).sum(1)[None, :]
    v, i = torch.topk(d2, 2, dim=1, largest=False)
    gap = (c[i[:, 1]] - c[i[:, 0]]).norm(dim=1)
    return i[:, 0], (v[:, 1] - v[:, 0]) / (2.0 * gap), c[i[:, 1]] - c[i[:, 0]]


def test_long_shapes_fp32_and_bf16_match_reference_golden(cuda_device, golden_dir):
    """2 layers at BASELINE shapes against the real reference's long fixture (T = 500 / 1500, -64 clamp, >1000 masked
    keys): fp32 within 1e-4 / tokens >= 99.5 %; bf16 within 1e-2 of the autocast reference run here."""
    g = np.load(os.path.join(golden_dir, 'conformer_long_l2.npz'))
    lengths = [int(v) for v in g['lengths']]
    wave, mask = _clips(lengths, int(g['total']))
    sd = synthetic_w2vbert_state_dict(2, 0)
    from audiotoken_b200.weights import synthetic_codebook
    cb = synthetic_codebook(2048, 1024, 4)
    m = torch.from_numpy(g['attention_mask'].astype(bool))
    for prec, tol in (('fp32', 1e-4), ('bf16', 3e-2)):
        enc = Wav2VecBertEncoder(device='cuda:0', precision=prec, n_layers=2, state_dict=sd, codebook=cb)
        toks, hid = enc(wave.to(cuda_device), mask.to(cuda_device), tap_layer=2)
        torch.cuda.synchronize()
        hid = hid.cpu()
        for i in range(len(lengths)):
            rows = g[f'rows_{i}']
            e = rel_err(hid[i, rows], torch.from_numpy(g[f'hidden_{i}']))
            assert e < tol, (prec, i, e)
        agree = float((toks[:, 0].cpu()[m].numpy() == g['tokens'][m.numpy()]).mean())
        print(f'long golden [{prec}]: token agreement {agree:.4f}')
        assert agree >= (0.995 if prec == 'fp32' else 0.97), (prec, agree)
        if prec == 'bf16':
            _, _, ref_bf = R.reference_embeddings(wave, mask, sd, 2, 'cuda:0', autocast=True)
            for i in range(len(lengths)):
                v = int(m[i].sum())
                e = rel_err(hid[i, :v], ref_bf[i, :v])
                assert e < 1e-2, (i, e)


def test_bf16_tokens_vs_reference_as_deployed(cuda_device):
    """19 layers, bf16, BASELINE configs[1] (64 x 10 s) plus a ragged batch padded to 30 s; data-derived codebook."""
    n_layers = 19
    sd = synthetic_w2vbert_state_dict(n_layers, 0)
    batches = [_clips([160000] * 64, 160000, first=100),
               _clips([480000, 333333, 250000, 160000, 90000, 32000], 480000, first=300)]
    model = R.build_model(sd, n_layers, 'cuda:0')
    ref32, refbf, valid = [], [], []
    for wave, mask in batches:
        e32, am, _ = R.reference_embeddings(wave, mask, sd, n_layers, 'cuda:0', autocast=False, model=model, tf32=False)
        ebf, _, _ = R.reference_embeddings(wave, mask, sd, n_layers, 'cuda:0', autocast=True, model=model, tf32=True)
        ref32.append(e32)
        refbf.append(ebf)
        valid.append(am.bool())
    del model
    torch.cuda.empty_cache()
    flat32 = torch.cat([e[v] for e, v in zip(ref32, valid)])
    flatbf = torch.cat([e[v] for e, v in zip(refbf, valid)])
    cb = data_derived_codebook(flat32, 2048, seed=2)
    enc = Wav2VecBertEncoder(device='cuda:0', precision='bf16', n_layers=n_layers, state_dict=sd, codebook=cb)
    mine_tok, mine_emb = [], []
    for (wave, mask), v in zip(batches, valid):
        toks, hid = enc(wave.to(cuda_device), mask.to(cuda_device), tap_layer=n_layers)
        torch.cuda.synchronize()
        mine_tok.append(toks[:, 0].cpu()[v].long())
        mine_emb.append(torch.nn.functional.layer_norm(hid.cpu(), (1024,))[v])
    mine_tok, mine_emb = torch.cat(mine_tok), torch.cat(mine_emb)

    tok32, margin, direction = _margins(flat32, cb)
    tokbf, _, _ = _margins(flatbf, cb)
    tok_dep = R.vq_eval_tokens(flatbf, cb, 'cuda:0', tf32=True)          # the deployed VQ (TF32 cdist) on ref-bf16
    noise = (flatbf.double() - flat32.double())
    sigma = float(noise.norm(dim=1).pow(2).mean().sqrt())                # RMS embedding noise of the reference's bf16 run
    sigma_proj = float(((noise * direction).sum(1) / direction.norm(dim=1)).pow(2).mean().sqrt())
    e_mine, e_ref = rel_err(mine_emb, flat32), rel_err(flatbf, flat32)
    ag = lambda a, b, sel=None: float((a == b).float().mean()) if sel is None else float((a[sel] == b[sel]).float().mean())  # noqa: E731
    i_, ii_, ip_ = ag(mine_tok, tokbf), ag(tokbf, tok32), ag(mine_tok, tok32)
    sel = margin > sigma
    sel4 = margin > 4.0 * sigma_proj
    iii_, iii4_ = ag(mine_tok, tok32, sel), ag(mine_tok, tok32, sel4)
    print(f'\\nbf16 vs reference as deployed, {flat32.shape[0]} valid rows, 19 layers, data-derived codebook 2048:\\n'
          f'  embeddings vs reference-fp32: this library {e_mine:.4f}, reference-bf16 (autocast) {e_ref:.4f}; '
          f'this library vs reference-bf16 {rel_err(mine_emb, flatbf):.4f}\\n'
          f'  (i)   this library vs reference-bf16 tokens      {i_:.4f}\\n'
          f'  (ii)  reference-bf16 vs reference-fp32 (floor)   {ii_:.4f}   [deployed TF32 VQ on ref-bf16 vs exact: {ag(tok_dep, tokbf):.4f}]\\n'
          f"  (i')  this library vs reference-fp32             {ip_:.4f}\\n"
          f'  (iii) rows with fp64 top-2 margin > RMS bf16 noise ({sigma:.3f}): {int(sel.sum())} rows, agreement {iii_:.4f}\\n'
          f'        rows with margin > 4 x projected noise ({sigma_proj:.4f}): {int(sel4.sum())} rows, agreement {iii4_:.4f}')
    assert e_mine <= 1.15 * e_ref + 1e-3, (e_mine, e_ref)
    assert int(sel4.sum()) > 0.5 * sel4.numel()
    assert iii4_ >= 0.995, iii4_
    if int(sel.sum()) >= 50:
        assert iii_ >= 0.995, iii_
    assert ip_ >= ii_ - 0.005, (ip_, ii_)
    assert i_ >= ii_ - 0.010, (i_, ii_)


def test_acoustic_bf16_codes_vs_reference_as_deployed(cuda_device):
    """All 16 codebooks: this library (tcgen05 bf16 encoder + exact RVQ) vs the EnCodec stand-in under CUDA autocast and
    in fp32.  Stage q is compared on the frames whose codes of all earlier stages agree (a flip changes every later
    residual), and against the reference's own bf16-vs-fp32 floor computed the same way."""
    from audiotoken_b200.acoustic import AcousticEncoder
    sd = synthetic_encodec_state_dict(0)
    wave = torch.stack([synthetic_waveform(400 + i, 10 * 24000, 24000) for i in range(16)])
    e32, c32 = R.acoustic_reference(wave, sd, 16, 'cuda:0', autocast=False, tf32=False)
    ebf, cbf = R.acoustic_reference(wave, sd, 16, 'cuda:0', autocast=True, tf32=True)
    enc = AcousticEncoder(device='cuda:0', state_dict=sd, precision='bf16')
    codes, emb = enc(wave.to(cuda_device), None, want_emb=True)
    torch.cuda.synchronize()
    assert enc.last_precision == 'bf16'
    codes = codes.cpu().long()
    e_mine, e_ref = rel_err(emb, e32), rel_err(ebf, e32)

    def staged(a, b):
        ok = torch.ones_like(a[:, 0], dtype=torch.bool)
        out = []
        for q in range(16):
            n = int(ok.sum())
            out.append(float((a[:, q][ok] == b[:, q][ok]).float().mean()) if n else float('nan'))
            ok &= a[:, q] == b[:, q]
        return out
    mine32, ref_floor, mine_bf = staged(codes, c32), staged(cbf, c32), staged(codes, cbf)
    print(f'\\nacoustic bf16 vs EnCodec stand-in as deployed (16 x 10 s): embeddings vs fp32: this library {e_mine:.4f}, '
          f'reference-bf16 {e_ref:.4f}')
    print('  stage-conditional agreement, codebooks 1..16')
    print('   this library vs ref-fp32 : ' + ' '.join(f'{v:.3f}' for v in mine32))
    print('   ref-bf16 vs ref-fp32     : ' + ' '.join(f'{v:.3f}' for v in ref_floor))
    print('   this library vs ref-bf16 : ' + ' '.join(f'{v:.3f}' for v in mine_bf))
    assert e_mine < 1e-2, e_mine
    assert e_mine <= 1.5 * e_ref + 1e-3, (e_mine, e_ref)
    for q in range(16):
        assert mine32[q] >= ref_floor[q] - 0.01, (q, mine32[q], ref_floor[q])
