"""Run one tcgen05 GEMM shape a few times (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, ops
M = int(os.environ.get('M', 32000)); N = int(os.environ.get('N', 4096)); K = int(os.environ.get('K', 1024))
epi = int(os.environ.get('EPI', L.EPI_BIAS_SWISH))
dev = torch.device('cuda:0')
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
W = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
bias = None if epi == L.EPI_GLU else torch.zeros(N, device=dev)
resid = torch.zeros(M, N, device=dev)
for _ in range(int(os.environ.get('ITERS', 6))):
    ops.gemm(A, W, bias, epi, 'bf16', L.IMPL_TENSOR, resid=resid)
torch.cuda.synchronize()
print('done')
