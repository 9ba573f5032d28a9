import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L
from audiotoken_b200.acoustic import AcousticEncoder
enc = AcousticEncoder(device='cuda:0', precision='bf16')
lib = L.load()
rows = 75 * 2000
emb = (torch.randn(rows, 128, device='cuda:0') * 0.9).contiguous()
def t(nq, dbg):
    lib.b2t_set_option(b'rvq_dbg', dbg)
    for _ in range(2): enc.rvq_encode(emb, L.IMPL_TENSOR, nq)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): enc.rvq_encode(emb, L.IMPL_TENSOR, nq)
    e.record(); torch.cuda.synchronize()
    print(f'n_q={nq:2d} dbg={dbg}: {s.elapsed_time(e)/3:.3f} ms', flush=True)
for nq in (1, 2, 16):
    t(nq, 0)
for dbg in (1, 2, 3, 4, 7):
    t(16, dbg)
