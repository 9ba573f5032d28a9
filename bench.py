#!/usr/bin/env python
"""Headline benchmark: audio-seconds tokenised per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the semantic_m encode path (log-mel front end -> 19 conformer layers ->
LayerNorm -> VQ 2048) over this rank's shard of BASELINE config[2]: 10 000 synthetic clips of
U(2, 30) s @16 kHz (seed 0), 1250 clips per GPU (weak scaling: per-GPU work is fixed), packed into
length-bucketed ragged batches.  Random-init weights of the named architecture (`data: synthetic`).

  value : whole-job audio-s/s with the waveforms already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public call with HOST (pinned) waveforms: per batch H2D copy, batch
          planning, encode, D2H copy of the int16 tokens — copies overlapped on a second stream
  roofline      : dominant kernel class = the tcgen05 GEMM; CUDA-event time of every GEMM launch of
                  instrumented steps vs MEASURED_PEAKS.json bf16 (sustained: timed inside a long step)
  cpu_baseline  : the oracle pipeline (oracle/, fp32 torch-CPU restatement of the reference) on the
                  host cores, on a bounded sample of the same workload, run the way the reference
                  runs it (clips padded to chunk_size = 30 s, all 21 layers of its checkpoint)

`--impl reference` prints the CPU arm alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
TOKEN_RATE = 50
TOTAL_CLIPS = 10000
CLIPS_PER_GPU = 1250
CHUNK_S = 30
N_LAYERS = 19
REF_LAYERS = 21
CODEBOOK = 2048
ROW_BUDGET = 65536          # token rows per ragged batch


def shard_lengths(rank: int, workload: str):
    """Clip lengths (samples) of this rank's shard."""
    if workload == 'c2':                      # BASELINE config[1] shape: 64 x 10 s
        return np.full(64, 10 * SR, dtype=np.int64)
    if workload == 'c4':                      # BASELINE config[3]: acoustic, 10k x 20 s @24 kHz, 1250 clips per GPU
        return np.full(CLIPS_PER_GPU, 20 * 24000, dtype=np.int64)
    g = torch.Generator().manual_seed(0)
    dur = torch.rand(TOTAL_CLIPS, generator=g, dtype=torch.float64) * 28.0 + 2.0
    lens = (dur * SR).round().to(torch.int64).numpy()
    r = rank % (TOTAL_CLIPS // CLIPS_PER_GPU)
    return lens[r * CLIPS_PER_GPU:(r + 1) * CLIPS_PER_GPU]


def synth_on_device(lengths: np.ndarray, seed: int, device, SR: int = SR) -> torch.Tensor:
    """Flat fp32 buffer of the clips back to back: 0.1*noise + 3 sinusoids per clip, in [-1, 1]."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = len(lengths)
    lens = torch.from_numpy(lengths).to(device)
    total = int(lengths.sum())
    x = 0.1 * torch.randn(total, generator=g, device=device)
    start = torch.cumsum(lens, 0) - lens
    t = (torch.arange(total, device=device) - torch.repeat_interleave(start, lens)).to(torch.float32) / SR
    for _ in range(3):
        f = torch.rand(n, generator=g, device=device) * (4000 - 80) + 80
        a = torch.rand(n, generator=g, device=device) * 0.15 + 0.05
        ph = torch.rand(n, generator=g, device=device) * 2 * math.pi
        x += torch.repeat_interleave(a, lens) * torch.sin(
            2 * math.pi * torch.repeat_interleave(f, lens) * t + torch.repeat_interleave(ph, lens))
    return x.clamp_(-1, 1)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, 100 ms period)."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown',
               0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting'}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1360.2), d.get('hbm_gbs', 6549.4), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(lengths: np.ndarray, steps: int, warmup: int, n_sample: int = 4):
    """Oracle pipeline on the host cores, run as the reference runs it (padded 30 s chunks, 21 layers)."""
    from audiotoken_b200.weights import synthetic_codebook, synthetic_w2vbert_state_dict, synthetic_waveform
    from oracle import conformer, fbank, quantize
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_w2vbert_state_dict(REF_LAYERS, seed=0)
    cb = synthetic_codebook(CODEBOOK, 1024, seed=4)
    lens = [int(v) for v in lengths[:n_sample]]
    pad = CHUNK_S * SR
    wave = torch.zeros(len(lens), pad)
    mask = torch.zeros(len(lens), pad)
    for i, n in enumerate(lens):
        wave[i, :n] = synthetic_waveform(i, n, SR)
        mask[i, :n] = 1
    audio_s = sum(lens) / SR

    def step():
        with torch.no_grad():
            feats, am = fbank.features(wave, mask)
            hs = conformer.hidden_states(feats, am, sd, REF_LAYERS)
            emb = conformer.final_embedding(hs[N_LAYERS])
            return quantize.kmeans_assign_fp32(emb.reshape(-1, 1024), cb)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return dict(value=audio_s / (ms / 1e3), unit='audio-s/s', cores=cores, kind='port',
                sample=f'{len(lens)} clips of the workload ({audio_s:.1f} audio-s) padded to {CHUNK_S} s, '
                       f'{REF_LAYERS} layers, fp32 torch-CPU oracle, {steps} timed steps'), ms


def cpu_reference_run_acoustic(steps: int, warmup: int):
    """BASELINE config[0]: EnCodec encode of one 10 s 24 kHz clip on the host CPU (oracle/seanet.py, fp32)."""
    from audiotoken_b200.weights import synthetic_encodec_state_dict, synthetic_waveform
    from oracle import seanet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_encodec_state_dict(0)
    wave = synthetic_waveform(0, 240000, 24000).unsqueeze(0)

    def step():
        with torch.no_grad():
            return seanet.rvq_codes_reference_fp32(seanet.encoder(wave, sd), sd, 16)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return dict(value=10.0 / (ms / 1e3), unit='audio-s/s', cores=cores, kind='port',
                sample=f'one 10 s 24 kHz clip (BASELINE configs[0]), SEANet + LSTM + RVQ-16, fp32 torch-CPU oracle, '
                       f'{steps} timed steps'), ms


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    lengths = shard_lengths(0, args.workload)
    if args.workload == 'c4':
        base, ms = cpu_reference_run_acoustic(max(1, args.steps), max(0, min(args.warmup, 1)))
    else:
        base, ms = cpu_reference_run(lengths, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {'impl': 'reference', 'metric': 'audio_seconds_per_second', 'value': base['value'], 'unit': 'audio-s/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, lengths), 'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def workload_config(workload, lengths):
    name = ('semantic_m encode_batch_files-equivalent: per-GPU shard of BASELINE configs[2] '
            f'({len(lengths)} of {TOTAL_CLIPS} synthetic clips, U(2,30) s @16 kHz, seed 0); '
            f'w2v-BERT 2.0 conformer x{N_LAYERS} + LayerNorm + VQ {CODEBOOK}x1024')
    if workload == 'c4':
        name = ('acoustic encode_batch_files-equivalent: per-GPU shard of BASELINE configs[3] (1250 of 10000 clips x 20 s '
                '@24 kHz); EnCodec SEANet encoder + LSTM + RVQ 16 codebooks')
    if workload == 'hs':
        name = ('semantic_s as the reference builds it (mHuBERT-base: 7-layer conv feature encoder + positional conv + 11 post-LN '
                'transformer layers + k-means 1000) on the per-GPU shard of BASELINE configs[2] (1250 clips, U(2,30) s @16 kHz)')
    if workload == 'c2':
        name = f'semantic_m encode of 64 x 10 s @16 kHz clips (BASELINE configs[1] shape); conformer x{N_LAYERS} + VQ {CODEBOOK}'
    return {'workload': name, 'clips_per_gpu': int(len(lengths)), 'audio_seconds_per_gpu': float(lengths.sum() / (24000 if workload == 'c4' else SR)),
            'row_budget_per_batch': (75 * 40000 if workload == 'c4' else ROW_BUDGET),
            'l2': 'inputs larger than L2: every batch streams >1 GB of activations through HBM'}


# ------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """process-wide state of the GPU arm: rank / device / collective helpers"""

    def __init__(self):
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        self.rank = int(os.environ.get('RANK', 0))
        self.local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(self.local)
        self.device = torch.device('cuda', self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_
            self.dist = dist_
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            self.dist.init_process_group('nccl', device_id=self.device)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def _reduce(self, v: float, op) -> float:
        if self.dist is None:
            return v
        t = torch.tensor([v], device=self.device, dtype=torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.SUM) if self.dist is not None else v

    def max_over_ranks(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.dist is not None else v

    def fault_check(self, where: str):
        """A device fault is asynchronous: surface it here, with a name, instead of in some tensor destructor."""
        from audiotoken_b200 import lib as L
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f'bench.py: rank {self.rank}: device fault detected after {where}: {e}\n'
                             f'  library says: {L.load().b2t_last_error().decode("utf-8", "replace")}\n'
                             f'  {L.device_trap_text() or "no device trap record"}\n'
                             '  re-run with B2T_DEBUG_SYNC=1 to name the kernel\n')
            sys.stderr.flush()
            os._exit(13)


def run_workload(ctx: Ctx, workload: str, steps: int, warmup: int, instrument: bool = True) -> dict:
    """One workload on this rank's shard: resident and end-to-end timing (+ per-kernel-class CUDA-event breakdown)."""
    from audiotoken_b200 import lib as L
    from audiotoken_b200 import packing
    from audiotoken_b200.encoder import Wav2VecBertEncoder
    device, rank = ctx.device, ctx.rank
    lengths = shard_lengths(rank, workload)
    acoustic = workload == 'c4'
    sr = 24000 if acoustic else SR
    audio_s = float(lengths.sum() / sr)
    if workload == 'hs':
        # the reference's own semantic_s (mHuBERT-base, hidden state 11, k-means 1000) on the c3 shard; raw waveforms in,
        # per-clip normalisation on the device, ceil(n / 320) token rows per clip as the reference saves them
        from audiotoken_b200.hubert import HubertEncoder, feat_lengths, plan_hubert
        enc = HubertEncoder(device=str(device), precision='bf16')
        t_pad = int(feat_lengths(CHUNK_S * sr))
        rows = np.minimum(np.array([packing.length_tokens(int(n), sr, TOKEN_RATE) for n in lengths]), t_pad)
        batches = packing.bucket_by_rows(rows.tolist(), 32768)

        def make_plan(ln, offs, rws):
            return plan_hubert(ln, offs, CHUNK_S * sr, rws)

        _plain = enc.encode_plan
        enc.encode_plan = lambda w, plan: _plain(w, plan, True)[:2]
    elif acoustic:
        from audiotoken_b200.acoustic import AcousticEncoder, plan_acoustic
        rows = np.array([packing.length_tokens(int(n), sr, 75) for n in lengths])
        enc = AcousticEncoder(device=str(device))
        batches = packing.bucket_by_rows(rows.tolist(), enc.max_rows_per_batch)

        def make_plan(ln, offs, rws):
            return plan_acoustic(ln, offs, np.maximum(rws * 320, ln), tiles=False)
    else:
        rows = np.array([packing.length_tokens(int(n), sr, TOKEN_RATE) for n in lengths])
        batches = packing.bucket_by_rows(rows.tolist(), ROW_BUDGET)
        enc = Wav2VecBertEncoder(device=str(device), precision='bf16', n_layers=N_LAYERS)

        def make_plan(ln, offs, rws):
            return packing.plan_semantic(ln, offs, CHUNK_S * sr, rws)

    # synthetic waveforms: generated on the device, kept there (value) and mirrored in pinned host memory (e2e)
    dev_waves, host_waves, plans, host_tokens = [], [], [], []
    for bi, idx in enumerate(batches):
        ln = lengths[idx]
        w = synth_on_device(ln, 1000 + rank * 1000 + bi, device, sr)
        offs = np.zeros(len(idx), dtype=np.int64)
        offs[1:] = np.cumsum(ln)[:-1]
        plan = make_plan(ln, offs, rows[idx])
        dev_waves.append(w)
        hw = torch.empty(w.numel(), dtype=torch.float32, pin_memory=True)
        hw.copy_(w)
        host_waves.append(hw)
        plans.append(plan)
        n_tok = plan.total_frames * enc.num_codebooks if acoustic else plan.total_rows
        host_tokens.append(torch.empty(n_tok, dtype=torch.int16, pin_memory=True))
    torch.cuda.synchronize()
    h2d_bytes = sum(h.numel() * 4 for h in host_waves)
    d2h_bytes = sum(h.numel() * 2 for h in host_tokens)
    launches = [0]

    def step_resident():
        for w, plan in zip(dev_waves, plans):
            enc.encode_plan(w, plan)
            launches[0] += enc.last_launches

    # end to end: uploads and token read-backs on their own streams (two copy engines), two staging buffers that alternate
    # over the batches of ALL steps — the upload of the next batch (of this step or the next one) runs under the encode of
    # the current one, as in the file loop (audiotoken_b200/core.py).  One stream for both directions would queue every
    # upload behind the previous batch's read-back, i.e. behind its encode: no overlap at all (acoustic: one batch per step).
    h2d_stream, d2h_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    max_samples = max(w.numel() for w in dev_waves)
    stage = [torch.empty(max_samples, dtype=torch.float32, device=device) for _ in range(2)]
    free_ev = [None, None]
    seq = [0]

    def step_e2e():
        comp = torch.cuda.current_stream()
        for idx, hw, ht in zip(batches, host_waves, host_tokens):
            k = seq[0] % 2
            seq[0] += 1
            buf = stage[k][:hw.numel()]
            with torch.cuda.stream(h2d_stream):
                if free_ev[k] is not None:
                    h2d_stream.wait_event(free_ev[k])                            # the encode that last read this buffer
                buf.copy_(hw, non_blocking=True)
                ready = h2d_stream.record_event()
            ln = lengths[idx]
            offs = np.zeros(len(idx), dtype=np.int64)
            offs[1:] = np.cumsum(ln)[:-1]
            plan = make_plan(ln, offs, rows[idx])                                # host planning is inside e2e
            comp.wait_event(ready)
            tokens, _ = enc.encode_plan(buf, plan)
            tokens = tokens.view(-1)
            done = comp.record_event()
            free_ev[k] = done
            tokens.record_stream(d2h_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                ht.copy_(tokens, non_blocking=True)
        comp.wait_stream(d2h_stream)            # the step's tokens are on the host before the step (and the timed region) ends

    def timed(fn, n):
        ctx.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        ctx.barrier()
        return ctx.max_over_ranks(s.elapsed_time(e) / n)

    for i in range(max(warmup, 3)):
        step_resident()
        ctx.fault_check(f'warm-up step {i} ({workload}, resident)')
    sampler = ClockSampler(ctx.local)
    sampler.start()
    launches[0] = 0
    ms_res = timed(step_resident, steps)
    gpu_launches = launches[0]
    clocks = sampler.result()
    ctx.fault_check(f'the timed resident steps ({workload})')
    step_e2e()
    ctx.fault_check(f'the e2e warm-up step ({workload})')
    ms_e2e = timed(step_e2e, steps)
    ctx.fault_check(f'the timed e2e steps ({workload})')

    res = dict(workload=workload, lengths=lengths, rows=rows, audio_s=audio_s, ms_res=ms_res, ms_e2e=ms_e2e,
               gpu_launches=gpu_launches, clocks=clocks, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes, acoustic=acoustic,
               num_codebooks=getattr(enc, 'num_codebooks', 1), n_batches=len(batches))
    if workload == 'hs':
        instrument = False
        res.update(cls_ms=np.zeros(6), gemm_flops=0.0, ac_ms=np.zeros(4))
    if instrument:          # instrumented steps: CUDA-event time per kernel class
        import ctypes as C
        lib = L.load()
        lib.b2t_profile_enable(1)
        cls_ms, gemm_flops, ac_ms = np.zeros(6), 0.0, np.zeros(4)
        for w, plan in zip(dev_waves, plans):
            enc.encode_plan(w, plan)
            if acoustic:
                arr4 = (C.c_float * 4)()
                L.check(lib.b2t_acoustic_profile_read(arr4), 'acoustic_profile_read')
                ac_ms += np.array(list(arr4))
            else:
                arr = (C.c_float * 6)()
                fl = C.c_double(0)
                L.check(lib.b2t_profile_read(arr, C.byref(fl)), 'profile_read')
                cls_ms += np.array(list(arr))
                gemm_flops += fl.value
        lib.b2t_profile_enable(0)
        res.update(cls_ms=cls_ms, gemm_flops=gemm_flops, ac_ms=ac_ms)
    del enc, dev_waves, host_waves, stage
    torch.cuda.empty_cache()
    return res


def run_files_e2e(ctx: Ctx, n_files: int) -> dict:
    """The public file loop, WAV files -> .npy on disk (SURVEY 8d's end-to-end definition): the first `n_files` clips of
    the shard are written as PCM16 WAV files (untimed), then AudioToken('semantic_m').encode_batch_files reads, decodes,
    encodes and writes them (timed with the wall clock around the call; one warm-up call on a small subset)."""
    import shutil
    import tempfile
    from audiotoken_b200 import AudioToken
    from audiotoken_b200 import io as aio
    lengths = shard_lengths(ctx.rank, 'c3')[:n_files]
    base = '/dev/shm' if os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK) else None
    root = tempfile.mkdtemp(prefix='b2t_files_', dir=base)
    try:
        indir, outdir = os.path.join(root, 'in'), os.path.join(root, 'out')
        os.makedirs(indir)
        wave = synth_on_device(lengths, 4242 + ctx.rank, ctx.device, SR).cpu()
        off = 0
        for i, n in enumerate(lengths):
            aio.write_wav(os.path.join(indir, f'clip{i:05d}.wav'), wave[off:off + int(n)], SR)
            off += int(n)
        in_bytes = sum(os.path.getsize(os.path.join(indir, f)) for f in os.listdir(indir))
        del wave
        tok = AudioToken('semantic_m', device=str(ctx.device), synthetic_weights=True, precision='bf16')
        tok.load_encoder()
        warm = sorted(os.path.join(indir, f) for f in os.listdir(indir))[:400]     # more than one window: allocator pools and pinned buffers at their working size
        tok.encode_batch_files(batch_size=ROW_BUDGET // 1500, outdir=os.path.join(root, 'warm'), chunk_size=CHUNK_S, audio_files=warm,
                               num_workers=min(16, os.cpu_count() or 1))
        ctx.barrier()
        t0 = time.perf_counter()
        tok.encode_batch_files(batch_size=ROW_BUDGET // 1500, outdir=outdir, chunk_size=CHUNK_S, audio_dir=indir,
                               num_workers=min(16, os.cpu_count() or 1))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = tok.last_stats
        out_bytes = sum(os.path.getsize(os.path.join(outdir, f)) for f in os.listdir(outdir))
        return {'value': st['audio_seconds'] / wall, 'unit': 'audio-s/s', 'wall_s': wall, 'files': st['files'],
                'audio_seconds': st['audio_seconds'], 'errors': len(st['errors']), 'windows': st['windows'], 'batches': st['batches'],
                'wav_bytes_read': int(in_bytes), 'npy_bytes_written': int(out_bytes), 'reader_threads': min(16, os.cpu_count() or 1),
                'host_seconds': {k: round(st[k], 3) for k in ('t_read_wait', 't_prepare', 't_launch', 't_fetch')},
                'what': f'AudioToken(semantic_m).encode_batch_files: {st["files"]} PCM16 WAV files on {"/dev/shm" if base else "tmp"} -> '
                        '.npy token files (file read + RIFF parse + GPU PCM decode + encode + D2H + npy write inside the timed region)'}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def load_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/r02_traffic.json)."""
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(p):
        return json.load(open(p))
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=['c3', 'c2', 'c4', 'hs'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the acoustic_c4 and files_e2e legs of the default run')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    from audiotoken_b200 import lib as L
    for kv in filter(None, os.environ.get('B2T_OPTS', '').split(',')):     # developer A/B switches, e.g. attn_two_pass=0
        k, v = kv.split('=')
        L.check(L.load().b2t_set_option(k.encode(), int(v)), kv)

    r = run_workload(ctx, args.workload, args.steps, args.warmup, instrument=True)
    acoustic, lengths, rows = r['acoustic'], r['lengths'], r['rows']
    ms_res, ms_e2e = r['ms_res'], r['ms_e2e']
    cls_ms, gemm_flops, ac_ms = r['cls_ms'], r['gemm_flops'], r['ac_ms']
    peak_tf, peak_hbm, peak_src = measured_peaks()
    ach = gemm_flops / (cls_ms[2] / 1e3) / 1e12 if cls_ms[2] > 0 else 0.0
    if acoustic:
        # 39.73 MFLOP per frame (encoder) + n_q * 2*1024*128 (RVQ)
        frames = float(rows.sum())
        flops = frames * (39.73e6 + r['num_codebooks'] * 2 * 1024 * 128)
        ach = flops / (ms_res / 1e3) / 1e12

    total_audio_s = ctx.sum_over_ranks(r['audio_s'])           # units all ranks processed per step
    extras = {}
    if world == 1 and not args.no_extras and args.workload == 'c3':
        # driver-visible lines for the other two measurement rows: BASELINE configs[3] on this GPU
        # and the public file loop, each a short pass after the main measurement
        a = run_workload(ctx, 'c4', min(args.steps, 3), 3, instrument=False)
        extras['acoustic_c4'] = {
            'value': a['audio_s'] / (a['ms_res'] / 1e3), 'unit': 'audio-s/s', 'ms_per_step': a['ms_res'],
            'e2e': {'value': a['audio_s'] / (a['ms_e2e'] / 1e3), 'unit': 'audio-s/s', 'ms_per_step': a['ms_e2e'],
                    'h2d_bytes_per_step': int(a['h2d_bytes']), 'd2h_bytes_per_step': int(a['d2h_bytes'])},
            'steps': min(args.steps, 3), 'gpu_launches': int(a['gpu_launches']), 'dtype': 'bf16',
            'config': workload_config('c4', a['lengths'])}
        extras['files_e2e'] = run_files_e2e(ctx, len(lengths))
    if rank == 0:
        value = total_audio_s / (ms_res / 1e3)
        e2e = total_audio_s / (ms_e2e / 1e3)
        traffic = load_traffic()
        n_gemm = 153 * r['n_batches']
        line = {
            'metric': 'audio_seconds_per_second', 'value': value, 'unit': 'audio-s/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_res, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': workload_config(args.workload, lengths), 'clocks': r['clocks'],
            'e2e': {'value': e2e, 'unit': 'audio-s/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': int(r['h2d_bytes']),
                    'd2h_bytes_per_step': int(r['d2h_bytes'])},
            'gpu_launches': int(r['gpu_launches']),
            'roofline': {'bound': 'tensor', 'kernel': 'gemm_tc_kernel (tcgen05 GEMM, all 153 launches per batch)',
                         'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                         'flops_per_launch': gemm_flops / n_gemm if n_gemm else None,
                         'ms_per_launch': float(cls_ms[2] / n_gemm) if n_gemm else None,
                         'traffic': (traffic or {}).get('gemm_tc_kernel', {}).get('dram_bytes_per_launch'),
                         'traffic_note': (traffic or {}).get('note'),
                         'peak_source': peak_src,
                         'gemm_share_of_step': float(cls_ms[2] / cls_ms.sum()) if cls_ms.sum() > 0 else None},
            'breakdown_ms_per_step': dict(zip(['fbank', 'layernorm', 'gemm', 'attention', 'dwconv', 'vq'],
                                              [float(v) for v in cls_ms])),
        }
        if traffic and not acoustic:
            line['roofline']['hbm_classes'] = traffic.get('hbm_classes')
        if acoustic:
            # per phase (CUDA events inside the library, one instrumented step): algorithmic FLOPs of SURVEY 8d
            # front end = 15 208 448 MAC/frame (18 convs minus the final one), LSTM 4 194 304, final conv 458 752
            ph_flops = [2 * 15208448.0 - 2 * 458752.0, 2 * 4194304.0, 2 * 458752.0, r['num_codebooks'] * 2.0 * 1024 * 128]
            phases = {}
            for nm, ms_p, fl in zip(['seanet_front_end', 'lstm', 'final_conv', 'rvq'], ac_ms, ph_flops):
                phases[nm] = {'ms': float(ms_p), 'tflops_algorithmic': (frames * fl / (ms_p / 1e3) / 1e12) if ms_p > 0 else None}
            # the strided-conv front end moves its intermediates through HBM: bytes the 11 kernels of a sub-batch read
            # and write per frame (bf16, channels-last, level 0 fused) = 286 976 B
            if ac_ms[0] > 0:
                phases['seanet_front_end']['hbm_gbps_moved'] = frames * 286976.0 / (ac_ms[0] / 1e3) / 1e9
                phases['seanet_front_end']['hbm_peak_gbps'] = peak_hbm
            line['roofline'].update({'kernel': 'whole acoustic step: seanet_conv0 + seanet_tc_kernel (tcgen05 conv/LSTM GEMMs) + '
                                               'rvq_tc_kernel (tcgen05 bf16x3), SURVEY 8d FLOPs / step time',
                                     'bound': 'tensor', 'gemm_share_of_step': None, 'phases': phases, 'traffic': None,
                                     'flops_per_launch': None, 'ms_per_launch': None})
            line['breakdown_ms_per_step'] = dict(zip(['seanet_front_end', 'lstm', 'final_conv', 'rvq'], [float(v) for v in ac_ms]))
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_reference_run_acoustic(3, 1) if acoustic else cpu_reference_run(lengths, steps=2, warmup=1)
            line['cpu_baseline'] = base
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == '__main__':
    main()
