"""Golden fixtures for the true `semantic_s` (mHuBERT-base) oracle — run in the build container only.

    python tests/golden/make_golden_hubert.py

Runs what reference audiotoken/encoder.py:60-108 runs, with the pieces that exist offline:
  * HF ``HubertModel(HubertConfig())`` (the architecture of `voidful/mhubert-base`, reference configs.py:50) loaded with
    this repo's seeded synthetic state dict instead of the checkpoint that cannot be downloaded here, called exactly as
    encoder.py:93: ``model.forward(input_batch, attention_mask=attention_mask, output_hidden_states=True)``;
  * HF ``Wav2Vec2FeatureExtractor()`` (what core.py:104 builds) for the waveform normalisation;
  * ``torch.nn.LayerNorm(768, elementwise_affine=False)``, ``torch.cdist``, ``argmin`` verbatim from encoder.py:75-81, 100-101;
  * a seeded 1000 x 768 stand-in for the joblib k-means centres (encoder.py:84-86).
Writes tests/golden/hubert.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from audiotoken_b200.weights import synthetic_hubert_state_dict, synthetic_waveform, synthetic_codebook  # noqa: E402


def main():
    from transformers import HubertConfig, HubertModel, Wav2Vec2FeatureExtractor
    torch.manual_seed(0)
    model = HubertModel(HubertConfig()).eval()
    sd = synthetic_hubert_state_dict(0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert set(missing) <= {'masked_spec_embed'}, missing
    fe = Wav2Vec2FeatureExtractor()
    lengths = [16000, 11111, 4000]
    total = 16000
    raw = [synthetic_waveform(300 + i, n, 16000) for i, n in enumerate(lengths)]
    norm = [fe(w.numpy(), sampling_rate=16000, return_tensors='pt').input_values[0] for w in raw]   # hubert_processor
    wave = torch.zeros(len(lengths), total)
    mask = torch.zeros(len(lengths), total)
    for i, (w, n) in enumerate(zip(norm, lengths)):
        wave[i, :n] = w                                     # datasets.py:99-103: zero right-padding + 0/1 mask
        mask[i, :n] = 1
    centres = synthetic_codebook(1000, 768, seed=9)
    ln = torch.nn.LayerNorm(768, elementwise_affine=False, bias=False).eval()
    with torch.no_grad():
        hs = model.forward(wave, attention_mask=mask, output_hidden_states=True).hidden_states
        emb = ln(hs[11])
        tok = torch.argmin(torch.cdist(emb, centres), dim=-1, keepdim=True).transpose(1, 2).to(torch.int16)
        feats = model.feature_extractor(wave).transpose(1, 2)
    np.savez_compressed(os.path.join(HERE, 'hubert.npz'), lengths=np.array(lengths), total=total,
                        norm0=norm[0].numpy()[:64], norm1_stats=np.array([float(norm[1].mean()), float(norm[1].std(unbiased=False))]),
                        feats=feats[:, ::7, ::16].numpy(), h0=hs[0][:, ::3].numpy(), h1=hs[1][:, ::3].numpy(),
                        h11=hs[11].numpy(), h12=hs[12][:, ::3].numpy(), tokens=tok.numpy())
    print('hubert.npz written: T =', hs[0].shape[1], 'tokens', tuple(tok.shape))


if __name__ == '__main__':
    main()
