"""Run the acoustic path a few times on a fixed batch (for ncu captures / timing)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
B = int(os.environ.get('B', 16)); SEC = float(os.environ.get('SEC', 20)); PREC = os.environ.get('PREC', 'fp32')
kw = {} if PREC == 'fp32' else {'precision': PREC}
enc = AcousticEncoder(device='cuda:0', **kw)
Ls = int(SEC * 24000)
wave = (0.1 * torch.randn(B * Ls, device='cuda:0')).clamp_(-1, 1)
plan = plan_acoustic([Ls] * B, np.arange(B) * Ls, [Ls] * B)
for i in range(int(os.environ.get('ITERS', 2))):
    torch.cuda.synchronize(); t0 = time.time()
    enc.encode_plan(wave, plan)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f'iter {i}: {dt*1e3:.1f} ms  {B*SEC/dt:.0f} audio-s/s  launches {enc.last_launches}', flush=True)
