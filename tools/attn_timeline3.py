"""Developer tool: clock64 timeline of one mid-grid CTA of the single-pass attention kernel (build with -DB2T_ATTN_TIMELINE)."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, packing
dev = torch.device('cuda:0')
lib = L.load()
lib.b2t_attention_set_dbg.argtypes = [C.c_void_p]
lib.b2t_set_option(b'attn_two_pass', 2)
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
    full = dbg.cpu().numpy()
    nkt = (rows_per_clip + 63) // 64
    sm = full[:256]; sm = sm[sm > 0]; t0 = sm[0]
    print(f'--- {n} x {rows_per_clip} rows, nkt = {nkt}: softmax thread (tile start, S ready, exps done, P buffer free, P handed over), relative clocks')
    for i in range(min(nkt, 12)):
        print('   ', (sm[5 * i:5 * i + 5] - t0).tolist())
    print('    loop end, O complete:', (sm[5 * nkt:5 * nkt + 2] - t0).tolist())
    si = full[256:384]; si = si[si > 0] - t0
    print('S issuer (before waits, K landed, S buffer free):')
    for j in range(min(nkt, 12)): print('   ', si[3 * j:3 * j + 3].tolist())
    pv = full[384:512]; pv = pv[pv > 0] - t0
    print('PV issuer (before waits, P ready, V landed):')
    for j in range(min(nkt, 12)): print('   ', pv[3 * j:3 * j + 3].tolist())
    tm = full[512:640]; tm = tm[tm > 0] - t0
    print('TMA K issue times:', tm[:12].tolist())
