"""Developer probe: nearest-centroid stage on the hidden states of production batches — time, re-scan count, bound."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from audiotoken_b200 import ops, packing
from audiotoken_b200.encoder import Wav2VecBertEncoder
from audiotoken_b200.weights import synthetic_codebook
dev = torch.device('cuda:0')
lengths = bench.shard_lengths(0, 'c3')
rows = np.array([packing.length_tokens(int(n), bench.SR, bench.TOKEN_RATE) for n in lengths])
batches = packing.bucket_by_rows(rows.tolist(), bench.ROW_BUDGET)
enc = Wav2VecBertEncoder(device='cuda:0', precision='bf16', n_layers=int(os.environ.get('LAYERS', 19)))
cb = synthetic_codebook(2048, 1024, 4).to(dev)
for bi in (0, 14):
    idx = batches[bi]
    ln = lengths[idx]
    offs = np.zeros(len(idx), dtype=np.int64); offs[1:] = np.cumsum(ln)[:-1]
    plan = packing.plan_semantic(ln, offs, bench.CHUNK_S * bench.SR, rows[idx])
    wave = bench.synth_on_device(ln, 1000 + bi, dev, bench.SR)
    tok, hid = enc.encode_plan(wave, plan, tap_layer=enc.n_layers)
    torch.cuda.synchronize()
    e = torch.nn.functional.layer_norm(hid, (1024,))
    sc = e @ cb.t() - 0.5 * (cb * cb).sum(1)[None, :]
    top = torch.topk(sc, 4, dim=1).values
    gap = (top[:, 0] - top[:, 3])
    print(f'batch {bi}: rows {plan.total_rows}; hidden finite {bool(torch.isfinite(hid).all())}; |hid| row-norm median {float(hid.norm(dim=1).median()):.1f}; '
          f'score top1-top4 gap: median {float(gap.median()):.3f} p01 {float(gap.quantile(0.01)):.4f} min {float(gap.min()):.5f}; '
          f'distinct tokens {int(torch.unique(tok).numel())}')
    stats = {}
    for it in range(3):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); o16, o32 = ops.vq_argmin(hid, cb, apply_ln=True, stats=stats if it == 2 else None); t.record(); torch.cuda.synchronize()
        print(f'   vq_argmin {s.elapsed_time(t):.3f} ms', stats)
    x = torch.randn_like(hid)
    for it in range(2):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.vq_argmin(x, cb, apply_ln=True, stats=stats); t.record(); torch.cuda.synchronize()
        print(f'   vq_argmin on randn rows {s.elapsed_time(t):.3f} ms', stats)
