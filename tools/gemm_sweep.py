"""Developer tool: time the tcgen05 GEMM on the conformer's layer shapes (M = 65536 rows), CUDA events, ITERS (20) launches
each after WARM (3) warm-up launches (environment; WARM=1 ITERS=1 under ncu --set full)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, ops
dev = torch.device('cuda:0')
M = int(os.environ.get('M', 65536))
ITERS, WARM = int(os.environ.get('ITERS', 20)), int(os.environ.get('WARM', 3))
shapes = (('ffn.w1 swish', 4096, 1024, L.EPI_BIAS_SWISH), ('ffn.w2', 1024, 4096, L.EPI_BIAS), ('qkv', 3072, 1024, L.EPI_BIAS),
          ('wo', 1024, 1024, L.EPI_BIAS), ('pw1 glu', 2048, 1024, L.EPI_GLU), ('pw2', 1024, 1024, L.EPI_BIAS))
tot = 0.0
for name, N, K, epi in shapes:
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = None if epi == L.EPI_GLU else torch.randn(N, device=dev) * 0.1
    for _ in range(WARM): ops.gemm(A, W, bias, epi, 'bf16')
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(ITERS): ops.gemm(A, W, bias, epi, 'bf16')
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / ITERS
    tot += ms * (2 if 'ffn' in name else 1)
    print(f'{name:14s} N={N} K={K}: {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:6.0f} TFLOP/s', flush=True)
print(f'layer total (2 x ffn): {tot * 1e3:.0f} us')
