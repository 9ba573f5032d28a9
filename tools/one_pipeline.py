"""Run the semantic pipeline a few times on a fixed batch (for ncu captures)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import packing
from audiotoken_b200.encoder import Wav2VecBertEncoder
B = int(os.environ.get('B', 64)); SEC = float(os.environ.get('SEC', 10)); LAYERS = int(os.environ.get('LAYERS', 2))
enc = Wav2VecBertEncoder(device='cuda:0', precision='bf16', n_layers=LAYERS)
Ls = int(SEC * 16000)
wave = (0.1 * torch.randn(B * Ls, device='cuda:0')).clamp_(-1, 1)
plan = packing.plan_semantic([Ls] * B, np.arange(B) * Ls, Ls)
for _ in range(int(os.environ.get('ITERS', 3))):
    enc.encode_plan(wave, plan)
torch.cuda.synchronize()
print('done', plan.total_rows)
