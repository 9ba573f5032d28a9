"""BASELINE configs[4]: nearest-centroid microbench, 1M frames x 1024-d vs K = 1024..16384 centroids (SURVEY 8d seeds).

For every K: time of b2t_vq_argmin (tensor fast pass + exact finalize), algorithmic FLOP rate 2*M*D*K / t against the
measured sustained bf16 peak (the kernel executes 3x that for the bf16x3 error compensation), rows that needed the
re-scan path, and bit-exactness against the fp64 oracle on a random sample of rows (and against the CUDA-core exact
kernel on the first 64k rows).  Prints one JSON line per K."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiotoken_b200 import lib as L, ops
from oracle import quantize

dev = torch.device('cuda:0')
M, D = int(os.environ.get('M', 1_000_000)), 1024
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops_sustained', 1360.2) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 1400.0
x = torch.randn(M, D, generator=torch.Generator().manual_seed(3)).to(dev)
sample = torch.randperm(M, generator=torch.Generator().manual_seed(5))[:2048]
for K in (1024, 2048, 4096, 8192, 16384):
    cb = torch.randn(K, D, generator=torch.Generator().manual_seed(4))
    cbd = cb.to(dev)
    ws = torch.empty(L.load().b2t_vq_workspace_bytes(M, D, K), dtype=torch.uint8, device=dev)
    stats = {}
    for _ in range(2):
        o16, o32 = ops.vq_argmin(x, cbd, workspace=ws, impl=L.IMPL_TENSOR)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        o16, o32 = ops.vq_argmin(x, cbd, workspace=ws, impl=L.IMPL_TENSOR)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    ops.vq_argmin(x, cbd, workspace=ws, impl=L.IMPL_TENSOR, stats=stats)
    idx, tie = quantize.nearest_centroid(x[sample].cpu(), cb)
    got = o32[sample].cpu().long()
    exact = bool(torch.equal(got[~tie], idx[~tie]))
    n_simt = 65536
    _, s32 = ops.vq_argmin(x[:n_simt].contiguous(), cbd, impl=L.IMPL_SIMT)
    same = bool(torch.equal(s32.cpu(), o32[:n_simt].cpu()))
    tf = 2.0 * M * D * K / ms / 1e9
    print(json.dumps({'workload': f'C5 nearest centroid {M} x {D} vs K={K}', 'ms': ms, 'frames_per_s': M / ms * 1e3,
                      'tflops_algorithmic': tf, 'tflops_executed_bf16x3': 3 * tf, 'frac_of_sustained_bf16_peak_executed': 3 * tf / peak,
                      'rescanned_rows': stats['n_fallback'], 'max_err_over_bound': stats['max_rel_err'] / 6.103515625e-5,
                      'oracle_sample_rows': int((~tie).sum()), 'bit_exact_vs_fp64_oracle': exact,
                      'equal_to_cuda_core_exact_kernel_first_64k': same}), flush=True)
