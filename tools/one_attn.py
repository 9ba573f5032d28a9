"""Launch the tcgen05 attention kernel a few times on a fixed ragged batch (for ncu captures).
ROWS / CLIPS select the shape (default 64 clips x 500 token rows = 64 x 10 s)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, packing
dev = torch.device('cuda:0')
lib = L.load()
rows = [int(os.environ.get('ROWS', 500))] * int(os.environ.get('CLIPS', 64))
lengths = [400 + 160 * (2 * r - 1) for r in rows]
offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
plan = packing.plan_semantic(lengths, offs, lengths, rows=rows)
M = plan.total_rows
qkv = (torch.randn(M, 3072, device=dev) * 0.7).to(torch.bfloat16)
E = torch.randn(73, 64, device=dev).to(torch.bfloat16)
db = packing.DeviceBatch(plan, dev); out = torch.zeros(M, 1024, device=dev, dtype=torch.bfloat16)
lib.b2t_set_option(b'attn_two_pass', int(os.environ.get('TWO_PASS', 1)))
for _ in range(int(os.environ.get('ITERS', 3))):
    L.check(lib.b2t_relkey_attention(qkv.data_ptr(), E.data_ptr(), db.byref(), out.data_ptr(), L.PREC_BF16, L.IMPL_TENSOR, L.stream_ptr()), 'attn')
torch.cuda.synchronize()
print('done', M, 'rows')
