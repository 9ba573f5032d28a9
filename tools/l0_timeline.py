"""Developer tool: per-tile timeline (clock64) of CTA 0 of the fused level-0 kernel."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
enc = AcousticEncoder(device='cuda:0', precision='bf16')
B, Ls = 16, 20 * 24000
wave = (0.1 * torch.randn(B * Ls, device='cuda:0')).clamp_(-1, 1)
plan = plan_acoustic([Ls] * B, np.arange(B) * Ls, [Ls] * B, tiles=False)
enc.encode_plan(wave, plan)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda:0')
enc.lib.b2t_seanet_set_l0_dbg.argtypes = [C.c_void_p]
enc.lib.b2t_seanet_set_l0_dbg(dbg.data_ptr())
enc.encode_plan(wave, plan)
torch.cuda.synchronize()
enc.lib.b2t_seanet_set_l0_dbg(None)
d = dbg.cpu().numpy().reshape(64, 8)
t0 = d[d > 0].min()
print('tile: b_start b_gotbuf b_done | mma1_issued | e1_start e1_done | e2_start e2_done   (clocks since start)')
for n in range(24):
    r = d[n] - t0
    print(f'{n:3d}: {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[7]:7d} | {r[3]:7d} {r[4]:7d} | {r[5]:7d} {r[6]:7d}')
print('steady-state period (clocks per tile, builder start):', (d[40, 0] - d[8, 0]) / 32)
