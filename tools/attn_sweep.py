"""Time the tcgen05 attention kernels alone (two-pass fixed-maximum vs online softmax vs mma.sync) on several clip lengths
and check the two tcgen05 variants against each other."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, ops, packing
dev = torch.device('cuda:0')
lib = L.load()
def plan_for(rows):
    lengths = [400 + 160 * (2 * r - 1) for r in rows]
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    return packing.plan_semantic(lengths, offs, lengths, rows=rows)
for name, rows in (('64x500', [500] * 64), ('24x1500', [1500] * 24), ('mix', list(np.random.default_rng(0).integers(100, 1500, 60))), ('short', [50, 64, 65, 1, 130, 200] * 20)):
    plan = plan_for([int(r) for r in rows])
    M = plan.total_rows
    qkv = (torch.randn(M, 3072, device=dev) * 0.7).to(torch.bfloat16)
    E = torch.randn(73, 64, device=dev).to(torch.bfloat16)
    flops = sum(4096.0 * r * r for r in rows)
    outs = {}
    for label, impl, two in (('persist-tmem', L.IMPL_TENSOR, 6), ('p-in-tmem', L.IMPL_TENSOR, 4), ('persistent', L.IMPL_TENSOR, 3), ('single-pass', L.IMPL_TENSOR, 2), ('two-pass', L.IMPL_TENSOR, 1), ('online', L.IMPL_TENSOR, 0), ('mma.sync', L.IMPL_MMA_SYNC, 0)):
        lib.b2t_set_option(b'attn_two_pass', two)
        db = packing.DeviceBatch(plan, dev); out = torch.zeros(M, 1024, device=dev, dtype=torch.bfloat16)
        for _ in range(2):
            L.check(lib.b2t_relkey_attention(qkv.data_ptr(), E.data_ptr(), db.byref(), out.data_ptr(), L.PREC_BF16, impl, L.stream_ptr()), 'attn')
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(10):
            L.check(lib.b2t_relkey_attention(qkv.data_ptr(), E.data_ptr(), db.byref(), out.data_ptr(), L.PREC_BF16, impl, L.stream_ptr()), 'attn')
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        outs[label] = out.float().cpu()
        print(f'{name} {label}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s', flush=True)
    ref = outs['online']
    for k in ('persist-tmem', 'p-in-tmem', 'persistent', 'single-pass', 'two-pass', 'mma.sync'):
        print(f'   {k} vs online: rel {float((outs[k] - ref).norm() / ref.norm()):.2e}  max-abs {float((outs[k] - ref).abs().max()):.2e}', flush=True)
lib.b2t_set_option(b'attn_two_pass', 4)
