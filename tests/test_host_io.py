"""CPU: on-disk token layout, segmentation rules, sharding (incl. a 2-rank gloo run)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from audiotoken_b200 import io as aio
from audiotoken_b200.configs import AudioConfig, Tokenizers, num_codebooks_to_bandwidth, AcousticEncoderConfig
from audiotoken_b200 import packing
from audiotoken_b200.sharding import lpt_shards, shard_files
from audiotoken_b200.weights import synthetic_waveform

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tokenizer_enum_and_configs():
    assert Tokenizers('semantic_m') is Tokenizers.semantic_m and str(Tokenizers.acoustic) == 'acoustic'
    assert [num_codebooks_to_bandwidth(k) for k in (2, 4, 8, 16)] == [1.5, 3, 6, 12]
    assert AcousticEncoderConfig(bandwidth=12).num_codebooks == 16
    assert AcousticEncoderConfig(bandwidth=1.5).num_codebooks == 2
    cfg = AudioConfig(file_name='a.wav', length_seconds=7.3, model_token_rate=50)
    assert cfg.length_tokens == math.ceil(7.3 * 50) == 365
    with pytest.raises(ValueError):
        AudioConfig(file_name='a.wav').length_tokens


def test_sanitize_path(tmp_path, monkeypatch):
    # reference test/utils.py: relative -> absolute, ~ expansion, mkdir
    monkeypatch.chdir(tmp_path)
    p = aio.sanitize_path('rel/dir')
    assert os.path.isabs(p) and os.path.isdir(p)
    monkeypatch.setenv('HOME', str(tmp_path))
    q = aio.sanitize_path('~/x/y')
    assert q == str((tmp_path / 'x' / 'y').resolve()) and os.path.isdir(q)


def test_token_file_layout_and_append(tmp_path):
    root = str(tmp_path)
    cfg = AudioConfig(file_name='/data/in/spk1/clip.v2.wav', length_seconds=1.0, model_token_rate=50)
    toks = torch.arange(2 * 60, dtype=torch.int16).reshape(2, 60)
    aio.save_audio_tokens(toks, cfg, root)
    path = os.path.join(root, 'clip.npy')                      # basename up to the FIRST '.'
    a = np.load(path)
    assert a.dtype == np.int16 and a.shape == (2, 50) and a.flags['C_CONTIGUOUS']
    aio.save_audio_tokens(toks + 1, cfg, root)                 # second chunk of the same file: appended on axis 1
    b = np.load(path)
    assert b.shape == (2, 100) and np.array_equal(b[:, :50], a) and np.array_equal(b[:, 50:], (toks + 1)[:, :50].numpy())
    aio.save_rel_audio_tokens(toks, cfg, root, '/data/in')
    assert np.load(os.path.join(root, 'spk1', 'clip.v2.npy')).shape == (2, 50)
    assert not [f for f in os.listdir(root) if '.tmp.' in f]


def test_wav_roundtrip_and_segments(tmp_path):
    sr = 16000
    x = torch.sin(torch.arange(sr * 7 + 1234) / 20.0) * 0.5
    p = str(tmp_path / 'a.wav')
    aio.write_wav(p, x, sr)
    assert aio.wav_info(p) == (sr, x.numel(), 1)
    y = aio.read_audio(p, sr)
    assert y.shape == (1, x.numel()) and float((y[0] - x).abs().max()) < 1e-4
    segs = list(aio.iter_segments(y, p, sr, 50, chunk_size=3))
    assert [s.wave.numel() for s in segs] == [48000, 48000, 16000 + 1234]      # 7 s + 1234 samples in 3 s chunks
    assert [s.config.length_tokens for s in segs] == [150, 150, math.ceil((16000 + 1234) / 16000 * 50)]
    # a trailing piece shorter than 3200 samples is skipped (reference datasets.py:95-97)
    short = list(aio.iter_segments(y[:, :48000 + 3199], p, sr, 50, chunk_size=3))
    assert len(short) == 1
    stereo = torch.stack([x, -x * 0.5])
    assert aio.convert_audio(stereo, sr, sr).shape == (1, x.numel())
    with pytest.raises(RuntimeError):
        aio.convert_audio(torch.zeros(3, 10), sr, sr)


def test_lpt_shards_balance_and_partition():
    rng = np.random.default_rng(0)
    dur = rng.uniform(2, 30, size=1000)
    for world in (1, 2, 4, 8):
        shards = lpt_shards(dur, world)
        assert sorted(i for s in shards for i in s) == list(range(1000))
        loads = [dur[s].sum() for s in shards]
        assert max(loads) - min(loads) <= 30.0
    files = [f'f{i}.wav' for i in range(10)]
    parts = [shard_files(files, list(range(10)), 3, r) for r in range(3)]
    assert sorted(f for p in parts for f in p) == files
    with pytest.raises(ValueError):
        shard_files(files, list(range(10)), 2, 2)


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 on CPU (gloo): both ranks derive the same disjoint partition and agree on the totals."""
    script = tmp_path / 'w.py'
    script.write_text(f'''
import os, sys, json
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
from audiotoken_b200.sharding import shard_files
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
files = [f"clip{{i:03d}}.wav" for i in range(101)]
dur = [2 + (i * 37 % 29) for i in range(101)]
mine = shard_files(files, dur, world, rank)
out = [None] * world
dist.all_gather_object(out, mine)
load = torch.tensor([float(sum(dur[files.index(f)] for f in mine))])
tot = load.clone(); dist.all_reduce(tot)
mx = load.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    allf = sorted(f for part in out for f in part)
    print(json.dumps(dict(ok=allf == sorted(files) and not set(out[0]) & set(out[1]), total=float(tot), max=float(mx))))
dist.destroy_process_group()
''')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29731', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    res = json.loads(line)
    assert res['ok'] and res['max'] <= res['total'] / 2 + 30


def test_acoustic_plan_tables():
    """Host planner of the acoustic path: level lengths (ceil division by the strides 2,4,5,8), the length-sorted
    LSTM order with its inverse, time-major prefix sums and the 320-sample alignment flag the tensor path needs."""
    from audiotoken_b200.acoustic import plan_acoustic
    lens = [3200, 320 * 7, 24000, 320]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    p = plan_acoustic(lens, offs, lens, tiles=False)
    assert p.aligned320 and p.total_frames == 10 + 7 + 75 + 1
    assert [int(v) for v in p.lens[4]] == [10, 7, 75, 1]
    assert [int(v) for v in p.lens[1]] == [1600, 1120, 12000, 160]
    assert list(p.order) == [2, 0, 1, 3] and list(p.rank[p.order]) == [0, 1, 2, 3]
    assert list(p.active[:2]) == [4, 3] and int(p.active[7]) == 2 and int(p.active[10]) == 1 and p.active.size == 75
    assert int(p.toff[-1]) == p.total_frames and list(p.toff[:3]) == [0, 4, 7]
    assert p.tile_clip[0].size == 0                      # tiles=False: no 64-row work lists
    q = plan_acoustic([333, 24137], [0, 333], [333, 24137])
    assert not q.aligned320 and [int(v) for v in q.lens[4]] == [2, 76] and q.tile_clip[0].size > 0
    with pytest.raises(ValueError):
        plan_acoustic([333], [0], [333], tiles=False)


class _FakeEncoder:
    """Host stand-in with the three calls encode_files makes; token j of a clip = (clip sum hash + j) mod 1000."""
    device = None
    max_rows_per_batch = 4000
    calls = 0

    def rows_for(self, padded_samples):
        return packing.padded_rows(padded_samples)

    def rows_for_tokens(self, n_tokens, padded_samples):
        return max(1, min(n_tokens, self.rows_for(padded_samples)))

    @staticmethod
    def expected(clip, rows):
        h = int(abs(float(clip.double().sum())) * 1000) % 997
        return ((h + np.arange(rows)) % 1000).astype(np.int16)[None, :]

    def encode_packed(self, clips, padded_samples, rows=None):
        type(self).calls += 1
        return [torch.from_numpy(self.expected(c, r)) for c, r in zip(clips, rows)]


def test_streaming_file_loop_windows_writes_and_errors(tmp_path, caplog):
    """encode_files with a host stand-in encoder: many small windows, every file written once with the tokens of its
    chunks in order, unreadable / unsupported files logged and skipped (reference datasets.py:136-137)."""
    from audiotoken_b200.core import encode_files
    sr, chunk = 16000, 2
    rng = np.random.default_rng(3)
    indir = tmp_path / 'in'
    indir.mkdir()
    files, lens = [], {}
    for i in range(23):
        n = int(rng.integers(3300, 5 * sr))
        p = indir / f'f{i:02d}.wav'
        aio.write_wav(str(p), synthetic_waveform(i, n, sr), sr)
        files.append(str(p))
        lens[str(p)] = n
    bad = indir / 'broken.wav'
    bad.write_bytes(b'RIFFnonsense')
    mp3 = indir / 'x.mp3'
    mp3.write_bytes(b'\\x00' * 64)
    files[5:5] = [str(bad), str(mp3)]
    out = tmp_path / 'out'
    _FakeEncoder.calls = 0
    with caplog.at_level('ERROR', logger='audiotoken_b200'):
        st = encode_files(_FakeEncoder(), files, aio.sanitize_path(out), sr, 50, chunk, batch_size=4, num_workers=3,
                          rel_dir=None, window_rows=700)
    assert st['files'] == 23 and set(st['errors']) == {str(bad), str(mp3)}
    assert st['windows'] > 3 and _FakeEncoder.calls >= st['windows']
    assert sum('skipping' in r.getMessage() for r in caplog.records) == 2
    for f in files:
        if f in st['errors']:
            continue
        got = np.load(aio.token_path_flat(f, str(out)))
        wave = aio.read_audio(f, sr)
        want = []
        for a in range(0, wave.shape[1], chunk * sr):
            seg = wave[0, a:a + chunk * sr]
            if seg.numel() < 3200:
                continue
            want.append(_FakeEncoder.expected(seg, packing.length_tokens(seg.numel(), sr, 50)))
        want = np.hstack(want)
        assert got.dtype == np.int16 and got.shape == want.shape and np.array_equal(got, want), f
    with pytest.raises(Exception):
        encode_files(_FakeEncoder(), [str(bad)], aio.sanitize_path(out), sr, 50, chunk, 4, 2, None, on_error='raise')


def test_chunk_output_lengths_match_the_host_resampler():
    """The streaming loop sizes one waveform buffer per window from chunk_output_lengths: it must predict exactly what
    convert_chunks returns (reference utils.py:82-101: chunks cut at the source rate, each resampled on its own)."""
    import numpy as np
    from audiotoken_b200 import io as aio
    for sr, n, chunk in [(16000, 16000 * 7 + 123, 3), (22050, 22050 * 4 + 17, 3), (8000, 8000 * 2 + 1, 1), (44100, 100000, 30),
                         (24000, 24000 * 5, 5), (48000, 48001, 1)]:
        pcm = (np.arange(n) % 200 - 100).astype(np.int16)
        got = [int(w.shape[-1]) for w in aio.convert_chunks(sr, pcm, 16000, chunk)]
        assert got == aio.chunk_output_lengths(sr, n, 16000, chunk), (sr, n, chunk, got)


def test_training_feature_stream_batching_on_a_host_stand_in(tmp_path):
    """iter_embeddings / train_codebook host logic without a GPU: segments are cut by the reference's chunk rules, batches
    respect the row budget, `padded_rows` switches between the rows that become tokens and the T rows per padded segment
    the reference trains on (cluster_tokens.py:83-136); the loop reports per batch and writes reference-named checkpoints."""
    import types
    import numpy as np
    import torch
    from audiotoken_b200 import io as aio
    from audiotoken_b200 import training
    from audiotoken_b200.packing import padded_rows
    from audiotoken_b200.weights import synthetic_waveform
    SR = 16000
    files = []
    for i, n in enumerate([52000, 16000, 33333, 90000]):
        p = tmp_path / f't{i}.wav'
        aio.write_wav(str(p), synthetic_waveform(i, n, SR), SR)
        files.append(str(p))

    class Enc:
        device = torch.device('cpu')
        config = types.SimpleNamespace(model_sample_rate=SR, model_token_rate=50)
        n_layers = 19
        seen = []

        def rows_for(self, pad):
            return padded_rows(pad)

        def rows_for_tokens(self, n_tokens, pad):
            return max(1, min(n_tokens, self.rows_for(pad)))

        def encode_plan(self, wave, plan, tap_layer=-1):
            assert tap_layer == 19
            self.seen.append(int(plan.total_rows))
            return None, torch.arange(plan.total_rows * 8, dtype=torch.float32).reshape(plan.total_rows, 8)

    enc = Enc()
    got = list(training.iter_embeddings(enc, files, chunk_size=5, max_rows=300))
    assert sum(b.shape[0] for b in got) == 163 + 50 + 105 + 250 + 32 and max(b.shape[0] for b in got) <= 300
    assert all(abs(float(b.mean())) < 1e-4 for b in got)                      # affine-free LayerNorm was applied
    enc.seen.clear()
    padded = list(training.iter_embeddings(enc, files, chunk_size=5, max_rows=600, padded_rows=True))
    assert sum(b.shape[0] for b in padded) == 5 * padded_rows(5 * SR) and enc.seen == [b.shape[0] for b in padded]

    class Trainer:
        codebook_size = 4
        calls = 0

        def train(self):
            return self

        def __call__(self, x):
            self.calls += 1
            return None, torch.arange(x.shape[0]) % 3, torch.tensor([0.25])

        def state_dict(self):
            return {'_codebook.embed': torch.zeros(1, 4, 8)}

    tr = Trainer()
    st = training.train_codebook(tr, iter(got), outdir=str(tmp_path / 'ck'), save_freq=2, layer=19)
    assert tr.calls == len(got) and st['batches'] == len(got) and st['active_fraction'] == 0.75 and st['commit_loss'] == 0.25
    assert sorted(os.listdir(tmp_path / 'ck'))[0] == 'quantizer__L19_C4_ckpt0.pkl'
