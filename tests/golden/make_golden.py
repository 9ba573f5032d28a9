"""Generate the golden fixtures from the REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_golden.py

Loads, read-only and without copying any source into this repo:
  * /root/reference/audiotoken/processors.py  (Wav2VecBertProcessor) under a stub ``audiotoken``
    package whose ``utils`` module holds only the three pure mel helpers, AST-extracted at run
    time from /root/reference/audiotoken/utils.py:286-328 (the real utils.py cannot be imported
    offline: torchaudio.io / datasets / network);
  * /root/reference/audiotoken/modeling_wav2vec2_bert.py (the SDPA relative-key patch), applied
    to HF ``Wav2Vec2BertSelfAttention`` exactly as reference encoder.py:14-15 does;
  * HF ``Wav2Vec2BertModel`` (what reference encoder.py:129 instantiates), loaded with this
    repo's seeded synthetic state dict (audiotoken_b200/weights.py) instead of the checkpoint
    that cannot be downloaded here;
  * HF ``EncodecModel(EncodecConfig())`` standing in for ``encodec.EncodecModel.encodec_model_24khz()``
    (reference encoder.py:38-52; the `encodec` package is not installable offline).

Writes small ``.npz`` files next to this script.  The GPU box never runs this script and never
reads /root/reference; tests only read the committed ``.npz`` files.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/audiotoken'
sys.path.insert(0, REPO)

from audiotoken_b200.weights import (synthetic_w2vbert_state_dict, synthetic_waveform,  # noqa: E402
                                     synthetic_codebook, synthetic_encodec_state_dict)


def load_reference_modules():
    src = open(os.path.join(REF, 'utils.py')).read()
    tree = ast.parse(src)
    wanted = {'hertz_to_mel', 'mel_to_hertz', 'create_triangular_filter_bank'}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert len(body) == 3
    utils = types.ModuleType('audiotoken.utils')
    utils.__dict__['torch'] = torch
    exec(compile(ast.Module(body=body, type_ignores=[]), 'ref_utils_extract', 'exec'), utils.__dict__)
    pkg = types.ModuleType('audiotoken')
    pkg.__path__ = [REF]
    sys.modules['audiotoken'] = pkg
    sys.modules['audiotoken.utils'] = utils
    spec = importlib.util.spec_from_file_location('audiotoken.processors', os.path.join(REF, 'processors.py'))
    proc = importlib.util.module_from_spec(spec)
    sys.modules['audiotoken.processors'] = proc
    spec.loader.exec_module(proc)
    spec = importlib.util.spec_from_file_location('audiotoken.modeling_wav2vec2_bert',
                                                  os.path.join(REF, 'modeling_wav2vec2_bert.py'))
    att = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(att)
    return proc, att


def make_clips(lengths, total, sr):
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        wave[i, :n] = synthetic_waveform(i, n, sr)
        mask[i, :n] = 1
    return wave, mask


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    proc_mod, att_mod = load_reference_modules()
    processor = proc_mod.Wav2VecBertProcessor()

    # ---- 1. front end: filters, window, features for ragged clips in one padded batch
    lengths = [16000, 11111, 8000, 3200]          # 1.0 s, odd length, 0.5 s, the 0.2 s minimum
    wave, mask = make_clips(lengths, 16000, 16000)
    with torch.no_grad():
        out = processor(wave, mask, 2)
    np.savez_compressed(os.path.join(HERE, 'fbank.npz'),
                        lengths=np.array(lengths), total=16000,
                        mel_filters=processor.mel_filters.detach().numpy(),
                        window=processor.window.detach().numpy(),
                        input_features=out['input_features'].numpy(),
                        attention_mask=out['attention_mask'].numpy())
    # unpadded single clips, as AudioToken.encode() feeds them (mask of ones, odd T -> 1 pad row)
    singles = {}
    for n in (4800, 16000 + 37):
        w, m = make_clips([n], n, 16000)
        with torch.no_grad():
            o = processor(w, m, 2)
        singles[f'feat_{n}'] = o['input_features'].numpy()
        singles[f'mask_{n}'] = o['attention_mask'].numpy()
    np.savez_compressed(os.path.join(HERE, 'fbank_single.npz'), **singles)

    # ---- 2. conformer: HF model + the reference's attention patch, synthetic weights
    from transformers import Wav2Vec2BertConfig, Wav2Vec2BertModel
    from transformers.models.wav2vec2_bert.modeling_wav2vec2_bert import Wav2Vec2BertSelfAttention
    Wav2Vec2BertSelfAttention.forward = att_mod.forward          # reference encoder.py:14-15
    for n_layers, tag in ((2, 'l2'), (19, 'l19')):
        sd = synthetic_w2vbert_state_dict(n_layers, seed=0)
        cfg = Wav2Vec2BertConfig(num_hidden_layers=n_layers)
        model = Wav2Vec2BertModel(cfg)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and set(missing) <= {'masked_spec_embed'}, (missing, unexpected)
        model.eval()
        with torch.no_grad():
            hs = model(out['input_features'], attention_mask=out['attention_mask'],
                       output_hidden_states=True).hidden_states
        assert len(hs) == n_layers + 1
        emb = torch.nn.functional.layer_norm(hs[n_layers], (1024,))          # encoder.py:175-176
        cb = synthetic_codebook(2048, 1024, seed=4)
        d = torch.cdist(emb.double(), cb.double().unsqueeze(0).expand(emb.shape[0], -1, -1))
        tok = torch.argmin(d, dim=-1)
        save = dict(tokens=tok.numpy().astype(np.int16),
                    hidden_last=hs[n_layers].numpy().astype(np.float32))
        if n_layers == 2:
            save['hidden_0'] = hs[0].numpy()
            save['hidden_1'] = hs[1].numpy()
        np.savez_compressed(os.path.join(HERE, f'conformer_{tag}.npz'), **save)
        print(tag, 'hidden', tuple(hs[n_layers].shape), 'tokens', tuple(tok.shape))
        del model


LONG_LENGTHS = [160000, 480000, 123456]        # 10 s (BASELINE configs[1]), 30 s (chunk_size, T = 1500), an odd length


def long_rows(valid_rows: int):
    """Rows of a clip kept in the long fixtures: the start, the rows around the left clamp (distance -64), a stride
    over the body and the last valid rows."""
    keep = set(range(0, 12)) | set(range(58, 72)) | set(range(0, valid_rows, 41)) | set(range(max(0, valid_rows - 12), valid_rows))
    return np.array(sorted(r for r in keep if r < valid_rows), dtype=np.int64)


def main_long():
    """2-layer goldens at BASELINE shapes (T = 500 and T = 1500 in one batch padded to 30 s): pins the -64 clamp of the
    relative-key bias (modeling_wav2vec2_bert.py:52), key masking over >1000 padded keys and long-row softmax to the
    real reference.  Only selected rows of the hidden state are stored (fp32), tokens for every valid row."""
    torch.manual_seed(0)
    torch.set_num_threads(16)
    proc_mod, att_mod = load_reference_modules()
    processor = proc_mod.Wav2VecBertProcessor()
    wave, mask = make_clips(LONG_LENGTHS, 480000, 16000)
    with torch.no_grad():
        out = processor(wave, mask, 2)
    from transformers import Wav2Vec2BertConfig, Wav2Vec2BertModel
    from transformers.models.wav2vec2_bert.modeling_wav2vec2_bert import Wav2Vec2BertSelfAttention
    Wav2Vec2BertSelfAttention.forward = att_mod.forward          # reference encoder.py:14-15
    n_layers = 2
    sd = synthetic_w2vbert_state_dict(n_layers, seed=0)
    model = Wav2Vec2BertModel(Wav2Vec2BertConfig(num_hidden_layers=n_layers))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {'masked_spec_embed'}
    model.eval()
    with torch.no_grad():
        hs = model(out['input_features'], attention_mask=out['attention_mask'], output_hidden_states=True).hidden_states
    emb = torch.nn.functional.layer_norm(hs[n_layers], (1024,))
    cb = synthetic_codebook(2048, 1024, seed=4)
    d = torch.cdist(emb.double(), cb.double().unsqueeze(0).expand(emb.shape[0], -1, -1))
    tok = torch.argmin(d, dim=-1)
    am = out['attention_mask']
    save = dict(lengths=np.array(LONG_LENGTHS), total=480000, tokens=tok.numpy().astype(np.int16),
                attention_mask=am.numpy().astype(np.uint8))
    for i in range(len(LONG_LENGTHS)):
        rows = long_rows(int(am[i].sum()))
        save[f'rows_{i}'] = rows
        save[f'hidden_{i}'] = hs[n_layers][i, rows].numpy().astype(np.float32)
        save[f'features_{i}'] = out['input_features'][i, rows].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'conformer_long_l2.npz'), **save)
    print('long', tuple(hs[n_layers].shape), {k: v.shape for k, v in save.items() if hasattr(v, 'shape')})


def main_acoustic():
    """EnCodec stand-in (HF EncodecModel) on the synthetic weights: embeddings + RVQ-16 codes."""
    from transformers import EncodecConfig, EncodecModel
    sd = synthetic_encodec_state_dict(0)
    model = EncodecModel(EncodecConfig())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith(('encoder', 'decoder')) or k.endswith('codebook.embed')]
    model.eval()
    save = {}
    # BASELINE config 1 shape scaled down (1 s), an odd length (right reflect extra padding), a very short clip
    for tag, lengths in (('a', [24000, 24000]), ('b', [24137]), ('c', [4800]), ('d', [333])):
        w = torch.stack([synthetic_waveform(20 + i, n, 24000) for i, n in enumerate(lengths)])
        with torch.no_grad():
            emb = model.encoder(w.unsqueeze(1))                            # reference encoder.py:48
            codes = model.quantizer.encode(emb, 12.0)                      # reference encoder.py:50-52 -> [16, B, T]
        save[f'lengths_{tag}'] = np.array(lengths)
        save[f'emb_{tag}'] = emb.numpy()
        save[f'codes_{tag}'] = codes.transpose(0, 1).numpy().astype(np.int16)   # [B, K, T] as encoder.py:54
        if tag in ('a', 'c'):
            with torch.no_grad():                                          # reference decoder.py:67-68
                deq = model.quantizer.decode(codes)
                wav = model.decoder(deq)
            save[f'deq_{tag}'] = deq.numpy()
            save[f'dec_{tag}'] = wav.numpy()
            print('acoustic decode', tag, tuple(deq.shape), tuple(wav.shape))
        print('acoustic', tag, tuple(emb.shape), tuple(codes.shape))
    np.savez_compressed(os.path.join(HERE, 'acoustic.npz'), **save)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'acoustic':
        main_acoustic()
    elif len(sys.argv) > 1 and sys.argv[1] == 'long':
        main_long()
    else:
        main()
        main_long()
        main_acoustic()
