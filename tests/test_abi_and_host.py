"""CPU: the C-ABI library loads and exports every declared symbol; host-side planner and tables."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from audiotoken_b200 import lib as L
from audiotoken_b200 import packing
from audiotoken_b200.fbank_tables import dense_mel_bank, povey_window, sparse_mel_bank
from oracle import fbank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, 'include')):
        if fn.endswith('.h'):
            src = open(os.path.join(ROOT, 'include', fn)).read()
            src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
            names |= set(re.findall(r'\b(b2t_[a-z0-9_]+)\s*\(', src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/*.h but not exported'
    assert set(L.EXPORTS) <= set(declared)
    assert lib.b2t_version() >= 100


def test_no_cpu_fallback():
    with pytest.raises(L.B2TError):
        L.require_device(torch.device('cpu'))
    if not torch.cuda.is_available():
        # a compute call without a GPU must fail loudly, not fall back
        lib = L.load()
        g = L.GemmArgs()
        a = torch.zeros(64, 64)
        g.A, g.W, g.out, g.lda, g.ldo = a.data_ptr(), a.data_ptr(), a.data_ptr(), 64, 64
        g.M = g.N = g.K = 64
        g.precision = L.PREC_FP32
        assert lib.b2t_gemm(ctypes.byref(g), None) != 0
        assert b'fallback' in lib.b2t_last_error() or b'CUDA' in lib.b2t_last_error()


def test_codebook_trainer_has_no_cpu_path():
    """The training step refuses a CPU device and, without a GPU, the ABI call itself fails."""
    from audiotoken_b200.training import CodebookTrainer
    with pytest.raises(L.B2TError):
        CodebookTrainer(64, 16, device='cpu')
    if not torch.cuda.is_available():
        lib = L.load()
        x = torch.zeros(8, 64)
        idx = torch.zeros(8, dtype=torch.int32)
        cb, avg, cs = torch.zeros(16, 64), torch.zeros(16, 64), torch.zeros(16)
        ws = torch.zeros(lib.b2t_vq_ema_workspace_bytes(8, 64, 16), dtype=torch.uint8)
        rc = lib.b2t_vq_ema_update(x.data_ptr(), 64, 8, 64, idx.data_ptr(), cb.data_ptr(), avg.data_ptr(), cs.data_ptr(), 16,
                                   0.8, 1e-5, 1.0, None, None, ws.data_ptr(), ws.numel(), None)
        assert rc != 0
        assert torch.equal(cb, torch.zeros(16, 64))


def test_product_tables_match_oracle():
    dense = dense_mel_bank()
    assert np.array_equal(dense.numpy(), fbank.mel_filters().numpy()[:256])
    assert np.array_equal(povey_window().numpy(), fbank.povey_window().numpy())
    start, count, weight = sparse_mel_bank(dense)
    rebuilt = np.zeros((256, 80), dtype=np.float32)
    for f in range(80):
        rebuilt[start[f]:start[f] + count[f], f] = weight[f, :count[f]]
    assert np.array_equal(rebuilt, dense.numpy())
    assert count.max() <= 32


@pytest.mark.parametrize('total', [16000, 16037, 4800, 480000])
def test_plan_matches_oracle_masks(total):
    rng = np.random.default_rng(total)
    lengths = [total] + [int(v) for v in rng.integers(3200, total + 1, size=5)]
    B = len(lengths)
    mask = torch.zeros(B, total)
    for i, n in enumerate(lengths):
        mask[i, :n] = 1
    nf = fbank.num_frames(total)
    fm = fbank.frame_mask(mask, nf)
    plan = packing.plan_semantic(lengths, np.arange(B) * total, total)
    # the oracle's attention mask = validity of sub-frame 0 after dropping an odd frame
    rem = nf % 2
    am = fm[:, :nf - rem].reshape(B, (nf - rem) // 2, 2)[:, :, 0]
    assert plan.valid_rows.tolist() == am.sum(1).long().tolist()
    T = packing.padded_rows(total)
    assert (plan.rows == T).all()
    assert plan.frame_off[-1] == sum(fbank.num_frames(n) for n in lengths)
    assert (plan.stack_frames <= np.diff(plan.frame_off)).all()
    # every row is covered exactly once by the conv tiles, every query row by the attention tiles
    for clips, starts, tile in ((plan.ctile_clip, plan.ctile_t0, packing.CTILE), (plan.qtile_clip, plan.qtile_q0, packing.QTILE)):
        cover = np.zeros(plan.total_rows, dtype=int)
        for c, s in zip(clips, starts):
            r0 = plan.row_off[c]
            cover[r0 + s:min(r0 + s + tile, plan.row_off[c + 1])] += 1
        assert (cover == 1).all()


def test_length_tokens_and_rows():
    # saved token count = ceil(len_s * 50) (reference configs.py:213-218) never exceeds T of a 30 s chunk
    for n in (3200, 116800, 160000, 479999, 480000):
        lt = packing.length_tokens(n, 16000, 50)
        assert lt == math.ceil(n / 16000 * 50)
        plan = packing.plan_semantic([n], [0], 480000, rows=[lt])
        assert plan.rows[0] == lt >= plan.valid_rows[0]
    with pytest.raises(ValueError):
        packing.plan_semantic([100], [0], 480000)


def test_bucket_by_rows():
    rows = [100, 1500, 700, 30, 1499, 800]
    b = packing.bucket_by_rows(rows, 2000)
    assert sorted(i for bb in b for i in bb) == list(range(6))
    assert all(sum(rows[i] for i in bb) <= 2000 or len(bb) == 1 for bb in b)
