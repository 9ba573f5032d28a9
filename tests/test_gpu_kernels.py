"""GPU (-m gpu): every kernel class through the C ABI against the CPU oracle / a plain fp32 torch
reference of the same op.  Tolerances: fp32 mode 1e-4 relative (north star), bf16 mode 1e-2."""
import os

import numpy as np
import pytest
import torch

from audiotoken_b200 import lib as L
from audiotoken_b200 import ops, packing
from audiotoken_b200.weights import synthetic_codebook, synthetic_w2vbert_state_dict, synthetic_waveform
from oracle import conformer, fbank, quantize

pytestmark = pytest.mark.gpu

ATTN_DEFAULT = 4          # library default of the attn_two_pass option (attention_tc.cu::g_attn_two_pass)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def bf(x):
    return x.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------- fbank
def _golden_clips(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fbank.npz'))
    lengths = [int(v) for v in g['lengths']]
    total = int(g['total'])
    return g, lengths, total


def test_fbank_matches_reference_golden(cuda_device, golden_dir):
    g, lengths, total = _golden_clips(golden_dir)
    B = len(lengths)
    wave = torch.zeros(B, total)
    for i, n in enumerate(lengths):
        wave[i, :n] = synthetic_waveform(i, n, 16000)
    plan = packing.plan_semantic(lengths, np.arange(B) * total, total)
    sd = synthetic_w2vbert_state_dict(0, seed=0)
    lw, lb = sd['feature_projection.layer_norm.weight'].to(cuda_device), sd['feature_projection.layer_norm.bias'].to(cuda_device)
    logmel, feats, ln_out, valid = ops.fbank_features(wave.to(cuda_device).view(-1), plan, lw, lb, 'fp32')
    T = plan.total_rows // B
    ref = torch.from_numpy(g['input_features'])
    got = feats.view(B, T, 160).cpu()
    assert got.shape == ref.shape
    # fp32 FFT rounding on the weakest mel bins (unit-variance features); the end-to-end bar is the hidden-state test
    assert float((got - ref).abs().max()) < 3e-4, float((got - ref).abs().max())
    assert np.array_equal(valid.view(B, T).cpu().numpy().astype(np.float32), g['attention_mask'])
    ln_ref = torch.nn.functional.layer_norm(ref, (160,), lw.cpu(), lb.cpu(), 1e-5)
    assert float((ln_out.view(B, T, 160).cpu() - ln_ref).abs().max()) < 5e-4


def test_fbank_packed_equals_padded_and_bf16_mel(cuda_device):
    # ragged clips packed back to back == the same clips in a padded batch (batch-composition invariance)
    lengths = [16000, 9999, 3200, 12345]
    clips = [synthetic_waveform(10 + i, n, 16000) for i, n in enumerate(lengths)]
    total = 16000
    padded = torch.zeros(len(lengths), total)
    for i, c in enumerate(clips):
        padded[i, :c.numel()] = c
    sd = synthetic_w2vbert_state_dict(0, seed=0)
    lw, lb = sd['feature_projection.layer_norm.weight'].to(cuda_device), sd['feature_projection.layer_norm.bias'].to(cuda_device)
    plan_a = packing.plan_semantic(lengths, np.arange(len(lengths)) * total, total)
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    plan_b = packing.plan_semantic(lengths, offs, total)
    fa = ops.fbank_features(padded.to(cuda_device).view(-1), plan_a, lw, lb, 'fp32')
    fb = ops.fbank_features(torch.cat(clips).to(cuda_device), plan_b, lw, lb, 'fp32')
    for a, b in zip(fa, fb):
        assert torch.equal(a, b)
    # oracle with the autocast cast point of the mel matmul (processors.py:184)
    mask = torch.zeros(len(lengths), total)
    for i, n in enumerate(lengths):
        mask[i, :n] = 1
    ref, _ = fbank.features(padded, mask, mel_in_bf16=True)
    got = ops.fbank_features(padded.to(cuda_device).view(-1), plan_a, lw, lb, 'fp32', mel_bf16=True)[1]
    d = (got.view(ref.shape).cpu() - ref).abs()
    # bf16 rounding of the mel energies can flip one ulp (2^-8 relative => ~4e-3 in ln) on a few bins
    assert float(d.mean()) < 2e-3 and float((d > 0.05).float().mean()) < 1e-3


# --------------------------------------------------------------------------------------- layernorm
@pytest.mark.parametrize('rows', [1, 37, 1000])
def test_layernorm(cuda_device, rows):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 1024, generator=g) * 3 + 0.5
    w, b = torch.randn(1024, generator=g), torch.randn(1024, generator=g)
    ref = torch.nn.functional.layer_norm(x, (1024,), w, b, 1e-5)
    xd, wd, bd = x.to(cuda_device), w.to(cuda_device), b.to(cuda_device)
    out = ops.layernorm(xd, wd, bd)
    assert float((out.cpu() - ref).abs().max()) < 2e-5
    out16 = ops.layernorm(xd, wd, bd, out_precision='bf16')
    assert torch.equal(out16.cpu(), out.cpu().to(torch.bfloat16)) or rel_err(out16.float(), ref) < 4e-3
    valid = (torch.arange(rows) % 3 != 0).to(torch.uint8).to(cuda_device)
    outz = ops.layernorm(xd, wd, bd, row_valid=valid)
    assert float(outz.cpu()[0::3].abs().max()) == 0.0
    noaff = ops.layernorm(xd, None, None)
    assert float((noaff.cpu() - torch.nn.functional.layer_norm(x, (1024,))).abs().max()) < 2e-5


# -------------------------------------------------------------------------------------------- GEMM
def _gemm_ref(A, W, bias, epi, alpha=1.0, resid=None, valid=None, bf16=False, round_resid=False):
    r = bf if bf16 else (lambda t: t)
    acc = A.double() @ W.double().t()
    v = r((acc + (bias.double() if bias is not None else 0)).float())
    if epi == L.EPI_BIAS:
        return v
    if epi == L.EPI_BIAS_SWISH:
        return r(v * torch.sigmoid(v))
    if epi == L.EPI_RESID:
        out = resid + alpha * v
        return r(out) if round_resid else out
    if epi == L.EPI_GLU:
        return r(v[:, 0::2] * torch.sigmoid(v[:, 1::2]))
    if epi == L.EPI_BIAS_MASK:
        return v * valid.float().unsqueeze(1)
    raise ValueError


GEMM_CASES = [(130, 1024, 160, L.EPI_BIAS_MASK), (257, 4096, 1024, L.EPI_BIAS_SWISH), (64, 1024, 4096, L.EPI_RESID),
              (300, 3072, 1024, L.EPI_BIAS), (129, 2048, 1024, L.EPI_GLU), (1, 1024, 1024, L.EPI_RESID)]


@pytest.mark.parametrize('M,N,K,epi', GEMM_CASES)
def test_gemm_simt_fp32(cuda_device, M, N, K, epi):
    g = torch.Generator().manual_seed(M * 7 + epi)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias = None if epi == L.EPI_GLU else torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g)
    valid = (torch.arange(M) % 4 != 1).to(torch.uint8)
    ref = _gemm_ref(A, W, bias, epi, 0.5, resid, valid)
    out = ops.gemm(A.to(cuda_device), W.to(cuda_device), None if bias is None else bias.to(cuda_device), epi, 'fp32',
                   L.IMPL_SIMT, resid=resid.clone().to(cuda_device), row_valid=valid.to(cuda_device), alpha=0.5)
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TENSOR])
@pytest.mark.parametrize('M,N,K,epi', GEMM_CASES + [(5000, 4096, 1024, L.EPI_BIAS_SWISH), (4099, 1024, 4096, L.EPI_RESID)])
def test_gemm_bf16(cuda_device, impl, M, N, K, epi):
    g = torch.Generator().manual_seed(M * 11 + epi)
    A, W = bf(torch.randn(M, K, generator=g)), bf(torch.randn(N, K, generator=g) * 0.05)
    bias = None if epi == L.EPI_GLU else bf(torch.randn(N, generator=g))
    resid = torch.randn(M, N, generator=g)
    valid = (torch.arange(M) % 4 != 1).to(torch.uint8)
    ref = _gemm_ref(A, W, bias, epi, 0.5, resid, valid, bf16=True, round_resid=True)
    out = ops.gemm(A.to(cuda_device, torch.bfloat16), W.to(cuda_device, torch.bfloat16),
                   None if bias is None else bias.to(cuda_device), epi, 'bf16', impl,
                   resid=resid.clone().to(cuda_device), row_valid=valid.to(cuda_device), alpha=0.5, round_resid=True)
    torch.cuda.synchronize()
    out = out.float().cpu()
    # identical bf16 inputs, fp32 accumulation: only summation order differs -> at most 1 bf16 ulp on a few elements
    assert rel_err(out, ref) < 2e-3, rel_err(out, ref)
    assert float(((out - ref).abs() > 0.02 * ref.abs() + 1e-2).float().mean()) < 1e-4


def test_gemm_tensor_equals_simt_bitwise_mostly(cuda_device):
    g = torch.Generator().manual_seed(99)
    M, N, K = 1000, 1024, 1024
    A = torch.randn(M, K, generator=g).to(cuda_device, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.03).to(cuda_device, torch.bfloat16)
    bias = bf(torch.randn(N, generator=g)).to(cuda_device)
    o1 = ops.gemm(A, W, bias, L.EPI_BIAS, 'bf16', L.IMPL_SIMT).float()
    o2 = ops.gemm(A, W, bias, L.EPI_BIAS, 'bf16', L.IMPL_TENSOR).float()
    torch.cuda.synchronize()
    assert float((o1 != o2).float().mean()) < 0.02          # rare 1-ulp flips from summation order
    assert rel_err(o2, o1) < 1e-3


# --------------------------------------------------------------------------------------- attention
def _attn_case(seed, lengths_rows, valid_rows):
    """Build a packed qkv and the per-clip oracle output."""
    g = torch.Generator().manual_seed(seed)
    M = sum(lengths_rows)
    qkv = torch.randn(M, 3072, generator=g) * 0.7
    E = torch.randn(73, 64, generator=g)
    return qkv, E


def _attn_oracle(qkv, E, rows, valid, emu):
    """oracle.conformer.rel_key_attention without the projections (identity weights trick is too
    large), restated on the q/k/v tensors directly with the same formula."""
    outs = []
    off = 0
    for T, tv in zip(rows, valid):
        blk = qkv[off:off + T]
        q, k, v = (blk[:, i * 1024:(i + 1) * 1024].view(T, 16, 64).transpose(0, 1) for i in range(3))
        pos = torch.arange(T)
        dist = (pos.view(1, -1) - pos.view(-1, 1)).clamp(-64, 8) + 64
        r = torch.einsum('hld,rd->hlr', q, E)
        if emu:
            r = bf(r)
        bias = torch.gather(r, 2, dist.view(1, T, T).expand(16, T, T)) / 8.0
        s = torch.einsum('hld,hrd->hlr', q, k) / 8.0 + bias
        s[:, :, tv:] = float('-inf')
        p = torch.softmax(s, dim=-1)
        if emu:
            p = bf(p)
        o = torch.einsum('hlr,hrd->hld', p, v)
        outs.append(o.transpose(0, 1).reshape(T, 1024))
        off += T
    return torch.cat(outs, 0)


def _attn_plan(rows, valid):
    # lengths chosen so that plan_semantic reproduces (rows, valid_rows): len = 400 + 160*(2*valid-1)
    lengths = [400 + 160 * (2 * v - 1) for v in valid]
    pad = [400 + 160 * (2 * r - 1) for r in rows]
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    plan = packing.plan_semantic(lengths, offs, pad, rows=rows)
    assert plan.valid_rows.tolist() == list(valid) and plan.rows.tolist() == list(rows)
    return plan


ATTN_CASES = [([50], [50]), ([200, 64, 130], [199, 20, 130]), ([500, 75], [499, 74]), ([1, 2, 3], [1, 1, 2])]


@pytest.mark.parametrize('rows,valid', ATTN_CASES)
def test_attention_simt_fp32(cuda_device, rows, valid):
    qkv, E = _attn_case(len(rows), rows, valid)
    plan = _attn_plan(rows, valid)
    ref = _attn_oracle(qkv, E, rows, valid, False)
    out = ops.relkey_attention(qkv.to(cuda_device), E.to(cuda_device), plan, 'fp32', L.IMPL_SIMT)
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)


@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TENSOR, L.IMPL_MMA_SYNC])
@pytest.mark.parametrize('rows,valid', ATTN_CASES + [([1500, 300, 129], [1499, 257, 128])])
def test_attention_bf16(cuda_device, impl, rows, valid):
    qkv, E = _attn_case(10 + len(rows), rows, valid)
    qkv, E = bf(qkv), bf(E)
    plan = _attn_plan(rows, valid)
    ref = bf(_attn_oracle(qkv, E, rows, valid, True))
    out = ops.relkey_attention(qkv.to(cuda_device, torch.bfloat16), E.to(cuda_device, torch.bfloat16), plan, 'bf16', impl)
    assert rel_err(out.float(), ref) < 1e-2, rel_err(out.float(), ref)


@pytest.mark.parametrize('mode', [1, 2, 3, 4, 6])
def test_attention_tensor_large_dynamic_range(cuda_device, mode):
    """Adversarial logits for the fixed-bound (mode 1: two-pass) and lazily rescaled (mode 2: single-pass) tcgen05
    kernels: q, k scaled so that raw scores span hundreds of log2 units, the largest raw score and the largest relative
    bias sit on DIFFERENT keys (the two-pass bound max raw + max bias is then tens of octaves above every real score),
    a bias spike on the far-left bucket, rows whose maximum grows late (forces the rescale path), masked keys holding
    the largest raw scores of all.  Checked against the fp64 formula."""
    lib = L.load()
    rows, valid = [700, 130, 64], [650, 130, 40]
    g = torch.Generator().manual_seed(99)
    M = sum(rows)
    qkv = torch.randn(M, 3072, generator=g) * 0.7
    qkv[:, :2048] *= 3.0                                     # q.k / 8 up to +-100: e^100 dynamic range within a row
    # keys late in the first clip get a large norm along one direction every query shares: the row maximum grows at the end
    qkv[:, 0:1024:64] += 2.0
    qkv[600:650, 1024:2048:64] += 9.0
    qkv[650:700, 1024:2048] *= 6.0                           # masked keys (>= valid_rows) with the largest raw scores
    E = torch.randn(73, 64, generator=g) * 0.5
    E[0] *= 8.0                                              # spike on the clamped far-left bucket (distance <= -64)
    E[72] *= 5.0                                             # and on the far-right one
    qkv, E = bf(qkv), bf(E)
    plan = _attn_plan(rows, valid)
    ref = _attn_oracle(qkv.double(), E.double(), rows, valid, False)
    assert torch.isfinite(ref).all()
    L.check(lib.b2t_set_option(b'attn_two_pass', mode), 'attn_two_pass')
    try:
        out = ops.relkey_attention(qkv.to(cuda_device, torch.bfloat16), E.to(cuda_device, torch.bfloat16), plan, 'bf16', L.IMPL_TENSOR)
    finally:
        L.check(lib.b2t_set_option(b'attn_two_pass', ATTN_DEFAULT), 'attn_two_pass')
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref.float()) < 1.5e-2, rel_err(out.float(), ref.float())


def test_attention_persistent_items(cuda_device):
    """The persistent single-pass kernels (attn_two_pass = 3: P in shared memory, 6: P in tensor memory) with their grid
    capped to 3 and to 7 CTAs: every CTA walks over dozens of (query tile, head) items of different lengths, so R pseudo
    tiles, ring counters and barrier phases carry across item boundaries.  Checked against the fp64 formula, and — with
    the polynomial exponential switched off — bit-for-bit against the one-item-per-CTA single-pass kernel (same
    arithmetic, same tiles; an integer bound makes the result independent of the rescaling schedule)."""
    lib = L.load()
    rows, valid = [300, 64, 130, 1, 257, 50, 128, 200], [290, 40, 130, 1, 257, 33, 128, 190]
    g = torch.Generator().manual_seed(5)
    qkv = bf(torch.randn(sum(rows), 3072, generator=g) * 0.9)
    E = bf(torch.randn(73, 64, generator=g) * 0.5)
    plan = _attn_plan(rows, valid)
    ref = _attn_oracle(qkv.double(), E.double(), rows, valid, False)
    outs = {}
    try:
        for mode, ctas, poly in ((2, 0, 0), (4, 0, 0), (3, 0, 0), (3, 3, 0), (3, 7, 0), (6, 0, 0), (6, 3, 0), (6, 7, 0),
                                 (4, 0, 1), (6, 0, 1), (6, 5, 1)):
            L.check(lib.b2t_set_option(b'attn_two_pass', mode), 'attn_two_pass')
            L.check(lib.b2t_set_option(b'attn_ctas', ctas), 'attn_ctas')
            L.check(lib.b2t_set_option(b'attn_poly_exp', poly), 'attn_poly_exp')
            outs[(mode, ctas, poly)] = ops.relkey_attention(qkv.to(cuda_device, torch.bfloat16), E.to(cuda_device, torch.bfloat16),
                                                            plan, 'bf16', L.IMPL_TENSOR).float().cpu()
    finally:
        L.check(lib.b2t_set_option(b'attn_two_pass', ATTN_DEFAULT), 'attn_two_pass')
        L.check(lib.b2t_set_option(b'attn_ctas', 0), 'attn_ctas')
        L.check(lib.b2t_set_option(b'attn_poly_exp', 1), 'attn_poly_exp')
    for k, o in outs.items():
        assert rel_err(o, ref.float()) < 1e-2, (k, rel_err(o, ref.float()))
    for k, o in outs.items():
        if k[2] == 0:
            assert torch.equal(o, outs[(2, 0, 0)]), k
        else:                               # degree-3 polynomial on a quarter of the exponentials: 1e-4 relative on P
            assert rel_err(o, outs[(2, 0, 0)]) < 2e-3, (k, rel_err(o, outs[(2, 0, 0)]))
            assert torch.equal(o, outs[(4, 0, 1)]), k


# ---------------------------------------------------------------------------------- depthwise conv
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_dwconv_ln_swish(cuda_device, precision):
    rows, valid = [70, 16, 33, 1], [70, 10, 33, 1]
    plan = _attn_plan(rows, valid)
    g = torch.Generator().manual_seed(7)
    M = sum(rows)
    x = torch.randn(M, 1024, generator=g)
    wd = torch.randn(1024, 31, generator=g) * 0.25
    lw, lb = torch.randn(1024, generator=g) * 0.1 + 1, torch.randn(1024, generator=g) * 0.1
    emu = precision == 'bf16'
    if emu:
        x = bf(x)
    outs, off = [], 0
    for T in rows:
        h = x[off:off + T].t().unsqueeze(0)
        hp = torch.nn.functional.pad(h, (30, 0))
        c = torch.nn.functional.conv1d(hp, (bf(wd) if emu else wd).unsqueeze(1), groups=1024)[0].t()
        if emu:
            c = bf(c)
        y = torch.nn.functional.layer_norm(c, (1024,), lw, lb, 1e-5)
        outs.append(y * torch.sigmoid(y))
        off += T
    ref = torch.cat(outs, 0)
    act = torch.bfloat16 if emu else torch.float32
    out = ops.dwconv_ln_swish(x.to(cuda_device, act), wd.t().contiguous().to(cuda_device), lw.to(cuda_device), lb.to(cuda_device), plan, precision)
    assert rel_err(out.float(), bf(ref) if emu else ref) < (1e-2 if emu else 2e-5)


@pytest.mark.parametrize('dc', [0.0, 6.0])
def test_dwconv_ring_variants_agree(cuda_device, dc):
    """The three bf16 implementations (direct loads, round-1 ring, round-2 ring with the one-pass LayerNorm tail) against an
    fp64 evaluation of the formula: a clip longer than several 64-row items (chained ring), clips that start inside a
    sub-tile's halo, a 1-row clip; `dc` adds a per-row offset of 6 standard deviations (mean^2 >> variance is the worst case
    for var = E[a^2] - mean^2)."""
    rows = [700, 16, 33, 1, 130, 64, 65]
    plan = _attn_plan(rows, rows)
    g = torch.Generator().manual_seed(11)
    M = sum(rows)
    x = bf(torch.randn(M, 1024, generator=g) + dc * torch.randn(M, 1, generator=g))
    wd = torch.randn(1024, 31, generator=g) * 0.25
    lw, lb = torch.randn(1024, generator=g) * 0.1 + 1, torch.randn(1024, generator=g) * 0.1
    outs, off = [], 0
    for T in rows:
        h = x[off:off + T].double().t().unsqueeze(0)
        c = torch.nn.functional.conv1d(torch.nn.functional.pad(h, (30, 0)), bf(wd).double().unsqueeze(1), groups=1024)[0].t()
        y = torch.nn.functional.layer_norm(bf(c.float()).double(), (1024,), lw.double(), lb.double(), 1e-5)
        outs.append(y * torch.sigmoid(y))
        off += T
    ref = torch.cat(outs, 0)
    lib = L.load()
    got = {}
    try:
        for mode in (0, 1, 2, 7, 8):
            L.check(lib.b2t_set_option(b'dwconv_ring', mode), 'dwconv_ring')
            got[mode] = ops.dwconv_ln_swish(x.to(cuda_device, torch.bfloat16), wd.t().contiguous().to(cuda_device), lw.to(cuda_device),
                                            lb.to(cuda_device), plan, 'bf16').double().cpu()
    finally:
        L.check(lib.b2t_set_option(b'dwconv_ring', 2), 'dwconv_ring')
    refb = bf(ref.float()).double()
    for mode, o in got.items():
        # an output is the bf16 rounding of a value within fp32 rounding of the formula — except where the fp32 conv sum sits
        # on a bf16 rounding boundary and the LayerNorm input moves by one ulp (a few outputs in 10^4)
        err = ((o - ref).abs() / (ref.abs() + 2e-2)).max().item()
        assert err < 4e-2, (mode, err)
        assert (o != refb).float().mean().item() < 2e-3, mode
    for mode in (0, 2, 7, 8):
        differ = (got[mode] != got[1]).float().mean().item()
        assert differ < 2e-3, (mode, differ)                     # one bf16 ulp on a few outputs at most
        assert (got[mode] - got[1]).abs().max().item() <= 4e-2 * max(1.0, ref.abs().max().item() / 4)


# ---------------------------------------------------------------------------------------------- VQ
@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TENSOR])
@pytest.mark.parametrize('M,D,K', [(1000, 1024, 2048), (777, 1024, 1000), (300, 128, 1024), (64, 256, 1), (50, 64, 3),
                                   (4000, 1024, 16384)])
def test_vq_argmin_bit_exact(cuda_device, M, D, K, impl):
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, D, generator=g)
    cb = torch.randn(K, D, generator=g)
    if K >= 1000:
        cb[K - 5:] = cb[:5]                                  # duplicated centroids: first index must win
        x[:5] = cb[:5] + 1e-3 * torch.randn(5, D, generator=g)
    idx, tie = quantize.nearest_centroid(x, cb)
    stats = {}
    o16, o32 = ops.vq_argmin(x.to(cuda_device), cb.to(cuda_device), impl=impl, stats=stats)
    torch.cuda.synchronize()
    assert torch.equal(o32.cpu().long()[~tie], idx[~tie])
    assert torch.equal(o16.cpu().long(), o32.cpu().long())
    # the certified error bound must hold with head-room: observed fast-pass error / bound scale
    bound = 6.103515625e-5 if impl == L.IMPL_TENSOR else 2.0 * D * 5.9604645e-8
    assert stats['max_rel_err'] < bound / 4, stats
    print(f'vq impl={impl} M={M} D={D} K={K}: re-scanned rows {stats["n_fallback"]}, '
          f'max observed err {stats["max_rel_err"]:.2e} (bound {bound:.2e})')


@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TENSOR])
def test_vq_argmin_near_ties_and_layernorm(cuda_device, impl):
    # rows placed (almost) on the bisector of two centroids: the fp64 re-check has to decide
    g = torch.Generator().manual_seed(5)
    cb = torch.randn(2048, 1024, generator=g)
    a, b = cb[torch.randint(0, 2048, (400,), generator=g)], cb[torch.randint(0, 2048, (400,), generator=g)]
    x = 0.5 * (a + b) + 1e-6 * torch.randn(400, 1024, generator=g)
    idx, tie = quantize.nearest_centroid(x, cb, tie_rel_margin=1e-12)
    _, o32 = ops.vq_argmin(x.to(cuda_device), cb.to(cuda_device), impl=impl)
    assert torch.equal(o32.cpu().long()[~tie], idx[~tie])
    # clusters of 4 nearly identical centroids: top-2 is not enough, the re-scan path must take over
    cb2 = cb.clone()
    for j in range(50):
        cb2[4 * j + 1:4 * j + 4] = cb2[4 * j] + 1e-5 * torch.randn(3, 1024, generator=g)
    x2 = cb2[0:200:4] + 0.05 * torch.randn(50, 1024, generator=g)
    idx2, tie2 = quantize.nearest_centroid(x2, cb2, tie_rel_margin=1e-13)
    stats = {}
    _, o2 = ops.vq_argmin(x2.to(cuda_device), cb2.to(cuda_device), impl=impl, stats=stats)
    assert torch.equal(o2.cpu().long()[~tie2], idx2[~tie2])
    assert stats['n_fallback'] >= 40, stats
    # fused affine-free LayerNorm (reference encoder.py:175-176)
    h = torch.randn(500, 1024, generator=g) * 2 + 1
    emb = conformer.final_embedding(h)
    idx3, tie3 = quantize.nearest_centroid(emb, cb)
    _, o = ops.vq_argmin(h.to(cuda_device), cb.to(cuda_device), apply_ln=True, impl=impl)
    assert (o.cpu().long()[~tie3] == idx3[~tie3]).float().mean() > 0.995


def test_vq_argmin_baseline_c5_full_size(cuda_device):
    """BASELINE configs[4] at full size (1M frames x 1024-d, K = 2048, SURVEY 8d seeds): the tensor fast pass +
    exact finalize against the fp64 oracle on a random sample of rows and against the CUDA-core exact kernel on a
    contiguous block; the size-independent property 'a centroid quantises to itself' on rows overwritten with centroids."""
    M, D, K = 1_000_000, 1024, 2048
    x = torch.randn(M, D, generator=torch.Generator().manual_seed(3))
    cb = torch.randn(K, D, generator=torch.Generator().manual_seed(4))
    x[-K:] = cb                                             # every centroid appears once as a frame
    xd, cbd = x.to(cuda_device), cb.to(cuda_device)
    stats = {}
    o16, o32 = ops.vq_argmin(xd, cbd, impl=L.IMPL_TENSOR, stats=stats)
    torch.cuda.synchronize()
    got = o32.cpu().long()
    assert torch.equal(got[-K:], torch.arange(K))
    sample = torch.randperm(M - K, generator=torch.Generator().manual_seed(5))[:1024]
    idx, tie = quantize.nearest_centroid(x[sample], cb)
    assert torch.equal(got[sample][~tie], idx[~tie])
    _, s32 = ops.vq_argmin(xd[:32768].contiguous(), cbd, impl=L.IMPL_SIMT)
    assert torch.equal(s32.cpu().long(), got[:32768])
    assert torch.equal(o16.cpu().long(), got)
    assert stats['max_rel_err'] < 6.103515625e-5 / 4, stats


def test_gemm_multicast_clusters_bit_identical(cuda_device):
    """2-CTA clusters with TMA multicast change data movement only: results equal the 1-CTA kernel bitwise."""
    lib = L.load()
    g = torch.Generator().manual_seed(123)
    for M, N, K, epi in [(5000, 4096, 1024, L.EPI_BIAS_SWISH), (4099, 1024, 4096, L.EPI_RESID), (300, 3072, 1024, L.EPI_BIAS),
                         (129, 2048, 1024, L.EPI_GLU), (257, 1024, 160, L.EPI_BIAS_MASK)]:
        A = torch.randn(M, K, generator=g).to(cuda_device, torch.bfloat16)
        W = (torch.randn(N, K, generator=g) * 0.05).to(cuda_device, torch.bfloat16)
        bias = None if epi == L.EPI_GLU else bf(torch.randn(N, generator=g)).to(cuda_device)
        resid = torch.randn(M, N, generator=g).to(cuda_device)
        valid = (torch.arange(M) % 4 != 1).to(torch.uint8).to(cuda_device)
        outs = []
        for mc in (0, 1):
            L.check(lib.b2t_set_option(b'gemm_multicast', mc), 'set_option')
            o = ops.gemm(A, W, bias, epi, 'bf16', L.IMPL_TENSOR, resid=resid.clone(), row_valid=valid, alpha=0.5)
            torch.cuda.synchronize()
            outs.append(o.float().cpu())
        L.check(lib.b2t_set_option(b'gemm_multicast', 1), 'set_option')
        assert torch.equal(outs[0], outs[1]), (M, N, K, epi)
