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
    ncta = plan.n_qtiles128 * 16 if hasattr(plan, 'n_qtiles128') else len(plan.qtile128_clip) * 16
    dbg = torch.zeros(1024 + 3 * ncta, dtype=torch.int64, device=dev)
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

    log = full[1024:].reshape(-1, 3)
    log = log[log[:, 1] > 0]
    busy, gaps, durs = 0, [], []
    for sm in np.unique(log[:, 0]):
        c = log[log[:, 0] == sm]; c = c[np.argsort(c[:, 1])]
        durs += (c[:, 2] - c[:, 1]).tolist()
        # two CTAs are resident at a time: CTA k starts when CTA k-2 (or earlier) has ended; gap = start - the latest end before it
        ends = np.sort(c[:, 2])
        for k in range(2, len(c)):
            prev = ends[ends <= c[k, 1]]
            if len(prev): gaps.append(int(c[k, 1] - prev[-1]))
        span = c[:, 2].max() - c[:, 1].min()
        busy += (c[:, 2] - c[:, 1]).sum() / (2.0 * span)
    print(f'CTA log: {len(log)} CTAs on {len(np.unique(log[:, 0]))} SMs; CTA duration median {np.median(durs):.0f} clocks (p10 {np.percentile(durs, 10):.0f}, p90 {np.percentile(durs, 90):.0f}); '
          f'gap from a slot freeing to the next CTA start: median {np.median(gaps):.0f}, p90 {np.percentile(gaps, 90):.0f}; mean slot occupancy {busy / len(np.unique(log[:, 0])):.3f}')
