"""Time the residual-VQ kernels alone (tensor vs CUDA-core) with CUDA events.  Developer tool."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L
from audiotoken_b200.acoustic import AcousticEncoder
enc = AcousticEncoder(device='cuda:0', precision='bf16')
rows = int(os.environ.get('ROWS', 75 * 4000))
emb = (torch.randn(rows, 128, device='cuda:0') * 0.9).contiguous()
for name, impl, iters in (('tensor', L.IMPL_TENSOR, 5), ('simt', L.IMPL_SIMT, 1)):
    for _ in range(2):
        codes = enc.rvq_encode(emb, impl)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc.rvq_stats()
    s.record()
    for _ in range(iters):
        codes = enc.rvq_encode(emb, impl)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    fl = 2.0 * rows * 128 * 1024 * 16
    print(f'rvq {name}: {rows} frames x 16 stages: {ms:.3f} ms  {rows/75/ms*1e3:.0f} audio-s/s  {fl/ms/1e9:.1f} TFLOP/s algorithmic'
          f'  stats {enc.rvq_stats()}', flush=True)
    if name == 'tensor':
        ref = codes.clone()
print('tensor == simt:', bool(torch.equal(ref, codes)))
