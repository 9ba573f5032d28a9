"""GPU (-m gpu): the codebook training step (b2t_vq_argmin + b2t_vq_ema_update) through the C ABI
against the fp64 oracle of the reference's `VectorQuantize(decay=0.8, commitment_weight=1)` training forward
(scripts/clustering/cluster_tokens.py:142-147, 293-311)."""
import pytest
import torch

from audiotoken_b200.training import CodebookTrainer
from oracle import quantize

pytestmark = pytest.mark.gpu


def _data(M, D, K, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(K, D, generator=g)
    x = centres[torch.randint(0, K, (M,), generator=g)] + 0.5 * torch.randn(M, D, generator=g)
    embed = centres + 0.3 * torch.randn(K, D, generator=g)
    return x, embed


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


@pytest.mark.parametrize('M,D,K', [(5000, 1024, 2048), (3333, 128, 1024), (700, 768, 1000)])
def test_ema_steps_match_oracle(cuda_device, M, D, K):
    """Three consecutive EMA steps (state carried on the device) == the fp64 oracle: indices bit-equal on every
    non-tie row, state within fp32 summation error, empty clusters decay, the checkpoint keeps the reference's keys."""
    tr = CodebookTrainer(D, K, decay=0.8, commitment_weight=1.0, device=cuda_device, return_quantized=True)
    x0, embed = _data(M, D, K, 7)
    sd = {'_codebook.embed': embed.unsqueeze(0), '_codebook.embed_avg': embed.unsqueeze(0).clone(),
          '_codebook.cluster_size': torch.ones(1, K), '_codebook.initted': torch.tensor([True])}
    tr.load_state_dict(sd)
    prev_avg, prev_cs = embed.clone(), torch.ones(K)
    for step in range(3):
        x, _ = _data(M, D, K, 7 + step)
        xd = x.to(cuda_device)
        old = tr.embed.clone()
        quant, idx, loss = tr(xd)
        torch.cuda.synchronize()
        # oracle step from the DEVICE state (so that assignment differences cannot accumulate across steps)
        ref_idx, ref_loss, e, avg, cs = quantize.vq_ema_train_step(x, old, prev_avg, prev_cs)
        _, tie = quantize.nearest_centroid(x, old)
        assert torch.equal(idx.cpu()[~tie], ref_idx[~tie])
        assert torch.equal(quant.cpu(), old.cpu()[idx.cpu()])
        assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
        assert _rel(tr.cluster_size, cs) < 1e-6
        assert _rel(tr.embed_avg, avg) < 1e-6
        assert _rel(tr.embed, e) < 1e-6
        prev_avg, prev_cs = tr.embed_avg.cpu().clone(), tr.cluster_size.cpu().clone()
    out = tr.state_dict()
    assert set(out) == {'_codebook.initted', '_codebook.cluster_size', '_codebook.embed_avg', '_codebook.embed'}
    assert tuple(out['_codebook.embed'].shape) == (1, K, D) and tuple(out['_codebook.cluster_size'].shape) == (1, K)


def test_ema_update_is_deterministic_and_handles_skew(cuda_device):
    """All rows on two centroids (one CTA sums 4000 rows), the rest empty; two runs are bit-identical."""
    D, K, M = 1024, 2048, 8000
    g = torch.Generator().manual_seed(3)
    embed = torch.randn(K, D, generator=g)
    x = torch.cat([embed[5] + 0.01 * torch.randn(M // 2, D, generator=g), embed[1999] + 0.01 * torch.randn(M // 2, D, generator=g)])
    x = x[torch.randperm(M, generator=g)]
    outs = []
    for _ in range(2):
        tr = CodebookTrainer(D, K, device=cuda_device)
        tr.load_state_dict({'_codebook.embed': embed.unsqueeze(0), '_codebook.cluster_size': torch.full((1, K), 2.0)})
        _, idx, loss = tr(x.to(cuda_device))
        torch.cuda.synchronize()
        outs.append((tr.embed.cpu().clone(), tr.embed_avg.cpu().clone(), tr.cluster_size.cpu().clone(), float(loss)))
    assert all(torch.equal(a, b) for a, b in zip(outs[0][:3], outs[1][:3])) and outs[0][3] == outs[1][3]
    ref_idx, ref_loss, e, avg, cs = quantize.vq_ema_train_step(x, embed, embed, torch.full((K,), 2.0))
    assert set(ref_idx.tolist()) == {5, 1999}
    assert _rel(outs[0][0], e) < 1e-6 and _rel(outs[0][2], cs) < 1e-6
    assert abs(outs[0][3] - ref_loss) <= 1e-5 * ref_loss


def test_eval_mode_leaves_state_untouched(cuda_device):
    tr = CodebookTrainer(128, 64, device=cuda_device).eval()
    g = torch.Generator().manual_seed(0)
    embed = torch.randn(64, 128, generator=g)
    tr.load_state_dict({'_codebook.embed': embed.unsqueeze(0)})
    x = torch.randn(2, 50, 128, generator=g)
    _, idx, _ = tr(x.to(cuda_device))
    assert idx.shape == (2, 50)
    assert torch.equal(tr.embed.cpu(), embed)
    ref, tie = quantize.nearest_centroid(x.reshape(-1, 128), embed)
    assert torch.equal(idx.cpu().reshape(-1)[~tie], ref[~tie])


def test_codebook_training_loop_over_files(cuda_device, tmp_path):
    """The clustering script's loop (cluster_tokens.py:83-136, 293-320) end to end on the device: WAV files -> LayerNormed
    hidden states of the tapped layer in ragged batches (iter_embeddings) -> one EMA step per batch (train_codebook) ->
    checkpoints with the reference's names and keys.  The streamed features equal the tap of a direct encoder call, and the
    loop leaves the same state as calling the trainer on those batches by hand."""
    import os
    from audiotoken_b200 import io as aio
    from audiotoken_b200.encoder import Wav2VecBertEncoder
    from audiotoken_b200.training import iter_embeddings, train_codebook
    from audiotoken_b200.packing import plan_semantic
    from audiotoken_b200.weights import synthetic_waveform
    SR = 16000
    files = []
    for i, n in enumerate([52000, 16000, 33333, 90000]):          # the last one spans two 5 s chunks
        p = tmp_path / f'train{i}.wav'
        aio.write_wav(str(p), synthetic_waveform(300 + i, n, SR), SR)
        files.append(str(p))
    enc = Wav2VecBertEncoder(device=cuda_device, precision='fp32', n_layers=2)
    batches = [b.clone() for b in iter_embeddings(enc, files, chunk_size=5, max_rows=300, layer=2)]
    rows = sum(b.shape[0] for b in batches)
    assert len(batches) >= 2 and rows == 163 + 50 + 105 + 250 + 32 and all(b.shape[1] == 1024 and b.is_cuda for b in batches)
    # one of the files, alone, through the operator: same LayerNormed rows somewhere in the stream
    wave = aio.read_audio_chunks(files[1], SR, 5, cuda_device)[0].reshape(-1)
    plan = plan_semantic([wave.numel()], [0], 5 * SR, [50])
    _, tap = enc.encode_plan(wave.contiguous(), plan, tap_layer=2)
    want = torch.nn.functional.layer_norm(tap, (1024,))
    allrows = torch.cat(batches)
    d = torch.cdist(want[:5].double(), allrows.double()).min(dim=1).values
    assert float(d.max()) < 1e-3 * float(want[:5].double().norm(dim=1).mean())

    K = 64
    g = torch.Generator().manual_seed(3)
    init = allrows[torch.randperm(rows, generator=g)[:K].to(allrows.device)].cpu()
    sd = {'_codebook.embed': init.unsqueeze(0), '_codebook.embed_avg': init.unsqueeze(0).clone(),
          '_codebook.cluster_size': torch.ones(1, K), '_codebook.initted': torch.tensor([True])}
    a = CodebookTrainer(1024, K, device=cuda_device); a.load_state_dict(sd)
    b = CodebookTrainer(1024, K, device=cuda_device); b.load_state_dict(sd)
    logs = []
    st = train_codebook(a, iter(batches), outdir=str(tmp_path / 'ckpt'), save_freq=2, layer=2, log=logs.append)
    for x in batches:
        b(x)
    assert st['batches'] == len(batches) and st['rows'] == rows and 0 < st['active_fraction'] <= 1 and len(logs) == len(batches)
    assert torch.equal(a.embed, b.embed) and torch.equal(a.cluster_size, b.cluster_size)
    assert not torch.equal(a.embed.cpu(), init)
    names = sorted(os.listdir(tmp_path / 'ckpt'))
    assert names[0] == f'quantizer__L2_C{K}_ckpt0.pkl' and len(names) == (len(batches) + 1) // 2
    ck = torch.load(tmp_path / 'ckpt' / names[-1])
    assert set(ck) == {'_codebook.initted', '_codebook.cluster_size', '_codebook.embed_avg', '_codebook.embed'}
    assert tuple(ck['_codebook.embed'].shape) == (1, K, 1024)
