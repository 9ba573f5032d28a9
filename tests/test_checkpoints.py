"""CPU: checkpoint readers (reference encoder.py:38, :132, :156-161) and the no-silent-synthetic-weights rule."""
import re

import pytest
import torch

from audiotoken_b200 import checkpoints as ck
from audiotoken_b200.weights import synthetic_codebook, synthetic_encodec_state_dict, synthetic_w2vbert_state_dict


def test_w2vbert_safetensors_roundtrip_and_prefix(tmp_path):
    from safetensors.torch import save_file
    sd = synthetic_w2vbert_state_dict(2, seed=0)
    d = tmp_path / 'w2vbert2_l21'
    d.mkdir()
    save_file({'wav2vec2_bert.' + k: v.contiguous() for k, v in sd.items()}, str(d / 'model.safetensors'))
    got = ck.load_w2vbert_state_dict(str(d))                      # the directory, as config.model_id is (configs.py:114-119)
    assert set(got) == set(sd) and ck.w2vbert_num_layers(got) == 2
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        ck.load_w2vbert_state_dict(str(tmp_path / 'nope'))


def test_vq_quantizer_pickle(tmp_path):
    cb = synthetic_codebook(2048, 1024, seed=4)
    p = tmp_path / 'quantizer.pkl'
    torch.save({'_codebook.initted': torch.tensor([1.0]), '_codebook.cluster_size': torch.ones(1, 2048),
                '_codebook.embed_avg': cb[None].clone(), '_codebook.embed': cb[None].clone()}, str(p))
    got = ck.load_vq_codebook(str(p))
    assert got.shape == (2048, 1024) and torch.equal(got, cb)


def test_kmeans_joblib(tmp_path):
    import joblib
    import numpy as np
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=4, n_init=1, random_state=0).fit(np.random.default_rng(0).normal(size=(64, 8)))
    p = tmp_path / 'km.bin'
    joblib.dump(km, str(p))
    got = ck.load_kmeans_centroids(str(p))
    assert got.shape == (4, 8) and np.allclose(got.numpy(), km.cluster_centers_.astype('float32'))


def test_encodec_package_names_map_to_hf(tmp_path):
    """HF-named synthetic tensors renamed the way the `encodec` package stores them, saved, loaded, mapped back."""
    sd = synthetic_encodec_state_dict(0)

    def to_pkg(k):
        k = re.sub(r'^(encoder|decoder)\.layers\.', r'\1.model.', k)
        k = k.replace('.parametrizations.weight.original0', '.weight_g').replace('.parametrizations.weight.original1', '.weight_v')
        m = re.match(r'^(decoder\.model\.(3|6|9|12))\.conv\.(.*)$', k)
        if m:                                                          # transposed convs
            return f'{m.group(1)}.convtr.convtr.{m.group(3)}'
        k = re.sub(r'\.conv\.(weight_g|weight_v|bias)$', r'.conv.conv.\1', k)
        return re.sub(r'^quantizer\.layers\.(\d+)\.codebook\.', r'quantizer.vq.layers.\1._codebook.', k)

    pkg = {to_pkg(k): v for k, v in sd.items()}
    assert 'encoder.model.0.conv.conv.weight_g' in pkg and 'quantizer.vq.layers.0._codebook.embed' in pkg
    assert 'decoder.model.3.convtr.convtr.weight_v' in pkg and 'encoder.model.13.lstm.weight_ih_l0' in pkg
    p = tmp_path / 'encodec_24khz.th'
    torch.save(pkg, str(p))
    got = ck.load_encodec_state_dict(str(p))
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)


def test_audiotoken_refuses_silent_synthetic_weights(monkeypatch):
    from audiotoken_b200 import AudioToken
    for v in (ck.ENV_W2VBERT, ck.ENV_VQ, ck.ENV_ENCODEC):
        monkeypatch.delenv(v, raising=False)
    for name in ('semantic_m', 'acoustic'):
        with pytest.raises(FileNotFoundError, match='synthetic_weights=True'):
            AudioToken(name, device='cuda:0').load_encoder()
