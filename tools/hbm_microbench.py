"""Time the HBM-bound kernels alone (dwconv+LN+swish, LayerNorm, fbank) with CUDA events.  Developer tool."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, packing
from audiotoken_b200.fbank_tables import DeviceFbankTables
dev = torch.device('cuda:0')
lib = L.load()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


B, rows = 128, 500          # 64 000 rows: 131 MB bf16 in + 131 MB out, larger than L2
lengths = [400 + 160 * (2 * rows - 1)] * B
offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
plan = packing.plan_semantic(lengths, offs, lengths)
db = packing.DeviceBatch(plan, dev)
M = plan.total_rows
s = L.stream_ptr()

x = torch.randn(M, 1024, device=dev).to(torch.bfloat16)
w = torch.randn(31, 1024, device=dev) * 0.1
g = torch.ones(1024, device=dev); b = torch.zeros(1024, device=dev)
out = torch.empty_like(x)
ms = timeit(lambda: L.check(lib.b2t_dwconv_ln_swish(x.data_ptr(), w.data_ptr(), g.data_ptr(), b.data_ptr(), db.byref(),
                                                    out.data_ptr(), L.PREC_BF16, s), 'dw'))
print(f'dwconv_ln_swish bf16 M={M}: {ms*1e3:.1f} us  {M*1024*4/ms/1e6:.0f} GB/s algorithmic', flush=True)

xf = torch.randn(M, 1024, device=dev)
o16 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: L.check(lib.b2t_layernorm(xf.data_ptr(), g.data_ptr(), b.data_ptr(), None, o16.data_ptr(), M, 1024,
                                              L.PREC_BF16, s), 'ln'))
print(f'layernorm f32->bf16 M={M}: {ms*1e3:.1f} us  {M*1024*6/ms/1e6:.0f} GB/s', flush=True)

wave = (torch.randn(sum(lengths), device=dev) * 0.1).clamp(-1, 1)
tabs = DeviceFbankTables(dev)
logmel = torch.empty(plan.total_frames, 80, device=dev)
ms = timeit(lambda: L.check(lib.b2t_fbank_logmel(wave.data_ptr(), db.byref(), tabs.byref(), logmel.data_ptr(), 0, s), 'fb'))
secs = sum(lengths) / 16000
print(f'fbank_logmel {secs:.0f} audio-s ({plan.total_frames} frames): {ms*1e3:.1f} us  '
      f'{(sum(lengths)*4 + plan.total_frames*320)/ms/1e6:.0f} GB/s algorithmic', flush=True)
