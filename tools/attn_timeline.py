"""Developer tool: clock64 timeline of one softmax warp of one mid-grid CTA of the two-pass attention kernel."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, packing
dev = torch.device('cuda:0')
lib = L.load()
lib.b2t_attention_set_dbg.argtypes = [C.c_void_p]
for rows_per_clip, n in ((500, 128), (1500, 48)):
    rows = [rows_per_clip] * n
    lengths = [400 + 160 * (2 * r - 1) for r in rows]
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    plan = packing.plan_semantic(lengths, offs, lengths, rows=rows)
    M = plan.total_rows
    qkv = (torch.randn(M, 3072, device=dev) * 0.7).to(torch.bfloat16)
    E = torch.randn(73, 64, device=dev).to(torch.bfloat16)
    db = packing.DeviceBatch(plan, dev); out = torch.zeros(M, 1024, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(640, dtype=torch.int64, device=dev)
    for it in range(2):
        lib.b2t_attention_set_dbg(dbg.data_ptr() if it else None)
        L.check(lib.b2t_relkey_attention(qkv.data_ptr(), E.data_ptr(), db.byref(), out.data_ptr(), L.PREC_BF16, L.IMPL_TENSOR, L.stream_ptr()), 'attn')
        torch.cuda.synchronize()
    lib.b2t_attention_set_dbg(None)
    full = dbg.cpu().numpy(); d = full[:128]; d = d[d > 0]; t0 = d[0]; d = d - t0
    mm = full[128:]; mm = mm[mm > 0] - t0
    nkt = (rows_per_clip + 63) // 64
    print(f'--- {n} x {rows_per_clip} rows, nkt = {nkt} ---')
    print('start, setup, R ready, R in smem:', d[:4].tolist())
    p1 = d[4:4 + nkt]
    print('pass-1 tile done:', p1.tolist())
    p2 = d[4 + nkt:4 + nkt + 3 * nkt].reshape(nkt, 3)
    print('pass-2 (S in regs, P buffer free, P handed over):')
    for i in range(nkt): print('   ', p2[i].tolist())
    print('O complete, stored, CTA done:', d[4 + 4 * nkt:].tolist())
    print('MMA warp, pass-1 S tiles (before KVFULL wait, K ready, S buffer free):')
    for j in range(nkt): print('   ', mm[3 * j:3 * j + 3].tolist())
    print('MMA warp, pass 2: S(i) [start, K ready, S free] then PV(i-1) [start, P ready, V ready]:')
    k = 3 * nkt
    for i in range(nkt):
        row = mm[k:k + 3].tolist(); k += 3
        if i > 0: row += mm[k:k + 3].tolist(); k += 3
        print('   ', row)
    print('    last PV:', mm[k:k + 3].tolist())
