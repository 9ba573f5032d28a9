"""GPU (-m gpu): production-size soak, determinism and multi-device tests.

Round 1's SCALE run aborted on the 8-GPU node: about one attention CTA in three million proceeded early (two
mbarrier.try_wait in flight in one warp) and produced a wrong 128-row tile or a launch failure.  Small parity tests
cannot see that; these tests run the kernels at the bench's batch geometry many times and require bit-identical
results, run two devices from one process, and run the two-rank launcher.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import bench
from audiotoken_b200 import lib as L
from audiotoken_b200 import ops, packing
from audiotoken_b200.encoder import Wav2VecBertEncoder

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_batches(rank=0):
    lengths = bench.shard_lengths(rank, 'c3')
    rows = np.array([packing.length_tokens(int(n), bench.SR, bench.TOKEN_RATE) for n in lengths])
    return lengths, rows, packing.bucket_by_rows(rows.tolist(), bench.ROW_BUDGET)


def _plan(lengths, rows, idx):
    ln = lengths[idx]
    offs = np.zeros(len(idx), dtype=np.int64)
    offs[1:] = np.cumsum(ln)[:-1]
    return ln, packing.plan_semantic(ln, offs, bench.CHUNK_S * bench.SR, rows[idx])


@pytest.mark.parametrize('batch', [14, 0])
def test_attention_is_deterministic_at_production_size(cuda_device, batch):
    """65 536-row ragged batches of the bench shard: batch 14 = 188 short clips (10 000 short-lived CTAs per launch, the
    geometry that exposed the fault), batch 0 = 44 long clips.  400 launches, all bit-identical."""
    lengths, rows, batches = _shard_batches()
    _, plan = _plan(lengths, rows, batches[batch])
    g = torch.Generator(device=cuda_device).manual_seed(7)
    qkv = (torch.randn(plan.total_rows, 3072, generator=g, device=cuda_device) * 0.7).to(torch.bfloat16)
    dist = (torch.randn(73, 64, generator=g, device=cuda_device) * 0.5).to(torch.bfloat16)
    ref = ops.relkey_attention(qkv, dist, plan, 'bf16')
    assert torch.isfinite(ref.float()).all()
    bad = 0
    for _ in range(400):
        bad += int(not torch.equal(ops.relkey_attention(qkv, dist, plan, 'bf16'), ref))
    assert bad == 0, f'{bad} of 400 launches deviated'


def test_pipeline_soak_320_production_batches(cuda_device):
    """20 passes over the 16 ragged batches of the bench shard (19 layers, bf16): every pass reproduces the tokens of
    the first one and no launch faults."""
    lengths, rows, batches = _shard_batches()
    enc = Wav2VecBertEncoder(device='cuda:0', precision='bf16', n_layers=bench.N_LAYERS)
    waves, plans = [], []
    for bi, idx in enumerate(batches):
        ln, plan = _plan(lengths, rows, idx)
        waves.append(bench.synth_on_device(ln, 1000 + bi, cuda_device, bench.SR))
        plans.append(plan)
    first = []
    for p in range(20):
        toks = [enc.encode_plan(w, plan)[0] for w, plan in zip(waves, plans)]
        torch.cuda.synchronize()
        if p == 0:
            first = [t.clone() for t in toks]
            for t in first:
                assert int(t.min()) >= 0 and int(t.max()) < 2048
        else:
            for bi, (a, b) in enumerate(zip(first, toks)):
                assert torch.equal(a, b), f'pass {p} batch {bi}: {int((a != b).sum())} tokens differ from pass 0'


def test_second_device_after_first_in_one_process():
    """cuda:1 after cuda:0 in one process: per-device shared-memory opt-ins, SM counts, twiddle tables and streams
    (they were process-wide statics).  Same weights and input => same tokens on both devices."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 visible GPUs')
    toks = []
    for d in (0, 1):
        dev = torch.device('cuda', d)
        with torch.cuda.device(dev):
            enc = Wav2VecBertEncoder(device=f'cuda:{d}', precision='bf16', n_layers=2)
            lens = np.array([48000, 31111, 16000], dtype=np.int64)
            wave = bench.synth_on_device(lens, 5, torch.device('cuda', 0), bench.SR).to(dev)
            offs = np.array([0, 48000, 79111], dtype=np.int64)
            plan = packing.plan_semantic(lens, offs, 48000)
            t, _ = enc.encode_plan(wave, plan)
            torch.cuda.synchronize(dev)
            toks.append(t.cpu())
    assert torch.equal(toks[0], toks[1])


def test_two_rank_launch_nccl(tmp_path):
    """`torchrun --nproc-per-node 2 bench.py --gpus 2`: one process per GPU, disjoint shards, max-over-ranks timing."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 visible GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29531', os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--steps', '1', '--warmup', '3',
           '--no-cpu-baseline']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    out = json.loads(line)
    assert out['n_gpus'] == 2 and out['value'] > 0 and out['e2e']['value'] > 0


def test_acoustic_is_deterministic_at_production_size(cuda_device):
    """600 x 20 s @24 kHz in one ragged batch (tcgen05 SEANet + LSTM step chain on two streams + tensor RVQ), 12 times:
    identical codes every time."""
    from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
    enc = AcousticEncoder(device='cuda:0', precision='bf16')
    n, length = 600, 20 * 24000
    lens = np.full(n, length, dtype=np.int64)
    wave = bench.synth_on_device(lens, 77, cuda_device, 24000)
    plan = plan_acoustic(lens, np.arange(n, dtype=np.int64) * length, lens, tiles=False)
    first = None
    for it in range(12):
        codes, _ = enc.encode_plan(wave, plan)
        torch.cuda.synchronize()
        assert enc.last_precision == 'bf16'
        if first is None:
            first = codes.clone()
            assert int(first.min()) >= 0 and int(first.max()) < 1024
        else:
            assert torch.equal(first, codes), f'iteration {it}: {int((first != codes).sum())} codes differ'


def test_device_trap_record_reaches_the_host():
    """The register-critical kernels (single-pass attention) report a protocol time-out by storing {site, a, b, block,
    thread} into mapped host memory before they trap (no printf ABI call inside setmaxnreg regions).  The trap poisons
    the CUDA context, so this runs in a child process: b2t_set_option('test_trap', 2) -> the synchronisation fails and
    b2t_last_device_trap returns the record the test kernel wrote."""
    code = (
        "import ctypes as C, sys\n"
        "sys.path.insert(0, %r)\n"
        "from audiotoken_b200 import lib as L\n"
        "import torch\n"
        "torch.zeros(1, device='cuda:0')\n"
        "lib = L.load()\n"
        "assert L.device_trap_text() == ''\n"
        "rc = lib.b2t_set_option(b'test_trap', 2)\n"
        "rec = (C.c_uint32 * 6)()\n"
        "got = lib.b2t_last_device_trap(rec)\n"
        "print('RC', rc, 'GOT', got, 'REC', list(rec), 'TEXT', L.device_trap_text(), 'ERR', lib.b2t_last_error().decode())\n"
        "import os; os._exit(0)\n" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    out = r.stdout
    assert 'RC' in out, (r.stdout[-2000:], r.stderr[-2000:])
    assert 'GOT 1' in out and 'REC [32343, 43981, 7, 0, 0, 0]' in out, out        # 0x7e57, 0xabcd, 7, block (0,0), thread 0
    assert 'RC 0' not in out                                                     # the launch's synchronisation reported the fault
    assert 'device trap record: site 0x7e57' in out
