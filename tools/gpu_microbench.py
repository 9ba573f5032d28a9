"""Per-kernel timings on one B200 (CUDA events, L2-flushing between iterations is not needed:
every working set here is larger than L2 or is timed as part of a chain).  Developer tool."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiotoken_b200 import lib as L, ops, packing  # noqa: E402
from audiotoken_b200.encoder import Wav2VecBertEncoder  # noqa: E402
from audiotoken_b200.weights import synthetic_waveform  # noqa: E402

dev = torch.device('cuda:0')


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    out = {}
    M = int(os.environ.get('M', 32000))
    shapes = [(4096, 1024, L.EPI_BIAS_SWISH), (1024, 4096, L.EPI_RESID), (3072, 1024, L.EPI_BIAS),
              (1024, 1024, L.EPI_RESID), (2048, 1024, L.EPI_GLU)]
    for N, K, epi in shapes:
        try:
            A = torch.randn(M, K, device=dev).to(torch.bfloat16)
            W = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
            bias = None if epi == L.EPI_GLU else torch.zeros(N, device=dev)
            resid = torch.zeros(M, N, device=dev)
            for mc in (0, 1):
                L.load().b2t_set_option(b'gemm_multicast', mc)
                ms = timeit(lambda: ops.gemm(A, W, bias, epi, 'bf16', L.IMPL_TENSOR, resid=resid), iters=20)
                out[f'gemm_tc_mc{mc}_N{N}_K{K}_e{epi}'] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
            # cuBLAS reference point (library; not used by the product)
            ms = timeit(lambda: torch.matmul(A, W.t()), iters=20)
            out[f'gemm_cublas_N{N}_K{K}'] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
        except Exception as ex:  # noqa: BLE001
            out[f'gemm_N{N}_K{K}_e{epi}'] = dict(error=str(ex))
        print(json.dumps({k: v for k, v in out.items() if f'N{N}_K{K}' in k}), flush=True)

    # whole pipeline, 64 x 10 s
    for prec, layers in (('bf16', 19),):
        try:
            enc = Wav2VecBertEncoder(device='cuda:0', precision=prec, n_layers=layers)
            B, Ls = 64, 160000
            wave = torch.stack([synthetic_waveform(i, Ls, 16000) for i in range(B)]).to(dev)
            lengths = [Ls] * B
            plan = packing.plan_semantic(lengths, np.arange(B) * Ls, Ls)
            flat = wave.view(-1)
            ms = timeit(lambda: enc.encode_plan(flat, plan), iters=5, warm=2)
            out[f'pipeline_{prec}_{layers}'] = dict(ms=ms, audio_s_per_s=B * 10.0 / ms * 1e3, launches=enc.last_launches)
            print(json.dumps({f'pipeline_{prec}_{layers}': out[f'pipeline_{prec}_{layers}']}), flush=True)
        except Exception as ex:  # noqa: BLE001
            print('pipeline error', ex, flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/microbench.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
