"""Run the acoustic path a few times on a fixed batch (for ncu captures / timing)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
B = int(os.environ.get('B', 16)); SEC = float(os.environ.get('SEC', 20)); PREC = os.environ.get('PREC', 'fp32')
kw = {} if PREC == 'fp32' else {'precision': PREC}
enc = AcousticEncoder(device='cuda:0', **kw)
if os.environ.get('L0F'):
    enc.lib.b2t_set_option(b'seanet_l0_fused', int(os.environ['L0F']))
if os.environ.get('PDL'):
    enc.lib.b2t_set_option(b'lstm_pdl', int(os.environ['PDL']))
if os.environ.get('SUB'):
    enc.lib.b2t_set_option(b'seanet_sub_frames', int(os.environ['SUB']))
Ls = int(SEC * 24000)
wave = (0.1 * torch.randn(B * Ls, device='cuda:0')).clamp_(-1, 1)
plan = plan_acoustic([Ls] * B, np.arange(B) * Ls, [Ls] * B)
import ctypes as C
enc.lib.b2t_profile_enable(int(os.environ.get('PROF', 0)))
for i in range(int(os.environ.get('ITERS', 2))):
    torch.cuda.synchronize(); t0 = time.time()
    enc.encode_plan(wave, plan)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms4 = (C.c_float * 4)(); enc.lib.b2t_acoustic_profile_read(ms4)
    if os.environ.get('PROF'): print('  phases ms: front %.1f lstm %.1f final %.1f rvq %.1f' % tuple(ms4))
    print(f'iter {i}: {dt*1e3:.1f} ms  {B*SEC/dt:.0f} audio-s/s  launches {enc.last_launches}', flush=True)
