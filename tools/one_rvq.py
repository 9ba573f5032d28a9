import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L
from audiotoken_b200.acoustic import AcousticEncoder
enc = AcousticEncoder(device='cuda:0', precision='bf16')
rows = int(os.environ.get('ROWS', 75 * 1000))
emb = (torch.randn(rows, 128, device='cuda:0') * 0.9).contiguous()
for _ in range(2):
    enc.rvq_encode(emb, L.IMPL_TENSOR)
torch.cuda.synchronize()
print('done', enc.rvq_stats())
