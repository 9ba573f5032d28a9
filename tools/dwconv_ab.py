"""A/B of the depthwise-conv + LayerNorm + swish kernels at production size (one ragged batch of ~65 k token rows):
time per launch (CUDA events, L2 flushed by the size of the tensors: 134 MB in + 134 MB out) and agreement of the
outputs with each other and with an fp64 evaluation of the same formula on the device.

    python tools/dwconv_ab.py [--dc 0|1]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, ops, packing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dc', type=float, default=0.0, help='per-row offset added to the conv input (stresses the one-pass variance)')
ap.add_argument('--iters', type=int, default=30)
ap.add_argument('--modes', default='0,1,2,7', help='dwconv_ring values; 4, 5, 6 are measurement-only variants (no tail / scalar, no tail / no taps)')
args = ap.parse_args()
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
secs = rng.uniform(2, 30, 200)
lengths, rws, tot = [], [], 0
for s_ in secs:
    n = int(s_ * 16000)
    r = packing.length_tokens(n, 16000, 50)
    if tot + r > 65536:
        break
    lengths.append(n)
    rws.append(r)
    tot += r
offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
plan = packing.plan_semantic(lengths, offs, [480000] * len(lengths), rows=rws)
M = int(plan.rows.sum())
g = torch.Generator(device=dev).manual_seed(1)
x = (torch.randn(M, 1024, device=dev, generator=g) + args.dc * torch.randn(M, 1, device=dev, generator=g)).to(torch.bfloat16)
wd = (torch.randn(31, 1024, device=dev, generator=g) * 0.25).contiguous()
lw = torch.randn(1024, device=dev, generator=g) * 0.1 + 1
lb = torch.randn(1024, device=dev, generator=g) * 0.1
lib = L.load()
db = ops.DeviceBatch(plan, dev)


def run(mode, out):
    L.check(lib.b2t_set_option(b'dwconv_ring', mode), 'set_option')
    L.check(lib.b2t_dwconv_ln_swish(x.data_ptr(), wd.data_ptr(), lw.data_ptr(), lb.data_ptr(), db.byref(), out.data_ptr(),
                                    L.PREC_BF16, L.stream_ptr()), 'dwconv')


outs = {}
for mode in [int(m) for m in args.modes.split(',')]:
    o = torch.empty_like(x)
    for _ in range(3):
        run(mode, o)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        run(mode, o)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / args.iters * 1e3
    print(f'dwconv_ring={mode}: {us:7.1f} us per launch, M = {M} rows, {M * 4096 / us / 1e3:6.0f} GB/s algorithmic '
          f'({M * 4096 / us / 1e3 / 6549.4:.2f} of the HBM copy peak), {2 * 31 * 1024 * M / us / 1e6:5.1f} TFLOP/s fp32', flush=True)
    if mode < 4 or mode >= 7:
        outs[mode] = o.clone()

# fp64 evaluation of the formula on a sample of clips (conv in fp64 of the bf16 inputs / bf16-rounded weights, rounded to bf16,
# LayerNorm and swish in fp64)
wq = wd.to(torch.bfloat16).double()
errs = {m: 0.0 for m in outs}
mism = {m: 0 for m in outs}
count = 0
row_off = np.concatenate([[0], np.cumsum(plan.rows)])
for ci in range(0, len(lengths), max(1, len(lengths) // 6)):
    r0, r1 = int(row_off[ci]), int(row_off[ci + 1])
    h = x[r0:r1].double().t().unsqueeze(0)
    c = torch.nn.functional.conv1d(torch.nn.functional.pad(h, (30, 0)), wq.t().unsqueeze(1), groups=1024)[0].t()
    c = c.to(torch.bfloat16).double()
    y = torch.nn.functional.layer_norm(c, (1024,), lw.double(), lb.double(), 1e-5)
    ref = (y * torch.sigmoid(y))
    refb = ref.to(torch.bfloat16)
    for m, o in outs.items():
        d = (o[r0:r1].double() - ref).abs()
        errs[m] = max(errs[m], float((d / (ref.abs() + 1e-2)).max()))
        mism[m] += int((o[r0:r1] != refb).sum())
    count += (r1 - r0) * 1024
for m in outs:
    print(f'dwconv_ring={m}: vs fp64 formula: max rel err {errs[m]:.2e}, {mism[m]} of {count} outputs differ from the bf16-rounded fp64 value '
          f'({mism[m] / count:.2e})')
for m in [m for m in outs if m != 1 and 1 in outs]:
    ne = outs[m] != outs[1]
    print(f'dwconv_ring={m} vs 1: {int(ne.sum())} of {ne.numel()} outputs differ ({float(ne.float().mean()):.2e}); max abs diff '
          f'{float((outs[m].float() - outs[1].float()).abs().max()):.3e}')
L.check(lib.b2t_set_option(b'dwconv_ring', 2), 'set_option')
