"""Developer tool: clock64 stamps of one softmax thread of the persistent P-in-TMEM attention kernel over four items
(build with -DB2T_ATTN_TIMELINE, tools/build_variants.py timeline:attention_tc.cu::-DB2T_ATTN_TIMELINE=1)."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, packing
dev = torch.device('cuda:0')
lib = L.load()
lib.b2t_attention_set_dbg.argtypes = [C.c_void_p]
lib.b2t_set_option(b'attn_two_pass', 6)
names = ['item start', 'S0 ready', 'S1 ready', 'S2 ready', 'loop end', 'quad barrier', 'next R taken', 'last PV retired', 'O stored', 'item done']
for rows_per_clip, n in ((500, 128), (1500, 48)):
    rows = [rows_per_clip] * n
    lengths = [400 + 160 * (2 * r - 1) for r in rows]
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    plan = packing.plan_semantic(lengths, offs, lengths, rows=rows)
    M = plan.total_rows
    qkv = (torch.randn(M, 3072, device=dev) * 0.7).to(torch.bfloat16)
    E = torch.randn(73, 64, device=dev).to(torch.bfloat16)
    db = packing.DeviceBatch(plan, dev); out = torch.zeros(M, 1024, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
    for it in range(2):
        lib.b2t_attention_set_dbg(dbg.data_ptr() if it else None)
        L.check(lib.b2t_relkey_attention(qkv.data_ptr(), E.data_ptr(), db.byref(), out.data_ptr(), L.PREC_BF16, L.IMPL_TENSOR, L.stream_ptr()), 'attn')
        torch.cuda.synchronize()
    lib.b2t_attention_set_dbg(None)
    t = dbg.cpu().numpy(); t = t[t > 0]
    nkt = (rows_per_clip + 63) // 64
    print(f'--- {n} x {rows_per_clip} rows, nkt = {nkt}; stamps relative to the first item start')
    for i in range(len(t) // 10):
        s = t[10 * i:10 * i + 10] - t[0]
        print('   ', ', '.join(f'{nm} {int(v)}' for nm, v in zip(names, s)))
