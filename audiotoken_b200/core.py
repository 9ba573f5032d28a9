"""Public API: ``AudioToken(tokenizer, device).encode / encode_batch_files``.

Same constructor and method signatures as the reference (audiotoken/core.py:27-289); the decode half
(audiotoken/core.py:291-359) is out of scope of this build.  Differences that do not change results:
  * weights are resolved lazily from plain paths (config fields or AUDIOTOKEN_* environment variables; nothing is
    downloaded); running without a checkpoint needs an explicit ``synthetic_weights=True`` (seeded synthetic weights
    of the named architecture: benchmark / parity mode);
  * ``encode_batch_files`` streams the corpus in bounded windows: reader threads prefetch files, each window is
    packed into ragged length-bucketed batches instead of padding every segment to ``chunk_size`` seconds, tokens
    come back once per batch on a side stream and a finished file's ``.npy`` is written at once, exactly once
    (atomic rename), by writer threads;
  * ``device`` must be an sm_100 CUDA device — there is no CPU path.
"""
from __future__ import annotations

import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import io as aio
from .configs import (AcousticEncoderConfig, AUDIO_EXTS, EncoderConfig, HubertEncoderConfig, SemanticSConfig, Tokenizers,
                      Wav2VecBertConfig, num_codebooks_to_bandwidth)
from .packing import bucket_by_rows, length_tokens, padded_rows

logger = logging.getLogger('audiotoken_b200')

# side streams of the file loop, one set per device for the life of the process: the caching allocator keeps a pool per
# stream, so fresh streams per call would meet cold pools (a cudaMalloc per window buffer) on every call
_SIDE_STREAMS: Dict[str, tuple] = {}


def _side_streams(device):
    """(token read-back, upload, ingest): uploads go on a copy-only stream (a pageable copy blocks the host until its
    stream has drained, so no kernel may sit in front of it), decode / resampling kernels on a high-priority stream (a
    few small CTAs that take the first SM the encoder's persistent kernels release)."""
    key = str(torch.device(device))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device),
                              torch.cuda.Stream(device=device, priority=-1))
    return _SIDE_STREAMS[key]


class AudioToken:
    def __init__(self, tokenizer: Union[Tokenizers, str], device: str = "cuda:0", compile: bool = False, **kwargs):
        self.tokenizer_name = Tokenizers(tokenizer)
        self.encoder: Optional[torch.nn.Module] = None
        self.decoder = None
        self.model_config: EncoderConfig
        self.transform_func = None
        self.compile = compile          # accepted for signature compatibility; kernels are ahead-of-time compiled
        self.kwargs = kwargs
        self.device = device
        self.num_codebooks = kwargs.get("num_codebooks", 16)
        assert self.num_codebooks in [2, 4, 8, 16], "num_codebooks must be one of [2, 4, 8, 16]"
        self.load_config()

    # reference core.py:73-90
    def load_config(self):
        if self.tokenizer_name == Tokenizers.acoustic:
            self.model_config = AcousticEncoderConfig(bandwidth=num_codebooks_to_bandwidth(self.num_codebooks))
        elif self.tokenizer_name == Tokenizers.semantic_s:
            # BASELINE.json defines semantic_s as a shallower w2v-BERT cut + k-means (the default here); the reference's
            # own semantic_s is mHuBERT-base + k-means (encoder.py:60-108): semantic_s_model='hubert'
            which = self.kwargs.get('semantic_s_model', os.environ.get('AUDIOTOKEN_SEMANTIC_S', 'w2vbert'))
            if which not in ('w2vbert', 'hubert'):
                raise ValueError("semantic_s_model must be 'w2vbert' or 'hubert'")
            self.model_config = HubertEncoderConfig() if which == 'hubert' else SemanticSConfig()
        elif self.tokenizer_name == Tokenizers.semantic_m:
            self.model_config = Wav2VecBertConfig()
        else:
            raise ValueError(f"Tokenizer {self.tokenizer_name} not supported")
        self.model_sample_rate = self.model_config.model_sample_rate

    # reference core.py:92-118
    def load_encoder(self):
        """Builds the encoder with the checkpoint the config (or AUDIOTOKEN_* environment) points at.  Explicit
        ``state_dict=`` / ``codebook=`` keyword arguments win.  Without any weights the encoders would tokenise with
        seeded synthetic weights of the named architecture: that is the benchmark / parity mode and has to be asked for
        with ``synthetic_weights=True`` — otherwise it raises (the reference downloads its checkpoints; there is no
        network here, so the files have to be provided)."""
        if self.encoder is not None:
            return
        from . import checkpoints as ck
        kw = dict(self.kwargs)
        enc_kw = {k: v for k, v in kw.items() if k in ('state_dict', 'codebook', 'precision', 'n_layers', 'seed')}
        synthetic_ok = bool(kw.get('synthetic_weights', False))
        cfg = self.model_config

        def need(what, env):
            if synthetic_ok:
                logger.warning('AudioToken(%s): no %s given, using SEEDED SYNTHETIC weights (tokens are meaningless; '
                               'benchmark / parity mode)', self.tokenizer_name, what)
                return
            raise FileNotFoundError(
                f'AudioToken({self.tokenizer_name}): no {what}. Set the config field, pass state_dict=/codebook=, or set ${env}; '
                'pass synthetic_weights=True to run with seeded synthetic weights (benchmarks / parity tests).')

        if self.tokenizer_name == Tokenizers.acoustic:
            from .acoustic import AcousticEncoder
            if 'state_dict' not in enc_kw:
                path = ck.resolve(getattr(cfg, 'weights', None), ck.ENV_ENCODEC)
                if path:
                    enc_kw['state_dict'] = ck.load_encodec_state_dict(path)
                else:
                    need('EnCodec 24 kHz checkpoint', ck.ENV_ENCODEC)
            enc_kw.pop('codebook', None)
            enc_kw.pop('n_layers', None)
            self.encoder = AcousticEncoder(config=cfg, device=self.device, **enc_kw)
            return
        hubert = isinstance(cfg, HubertEncoderConfig)
        if hubert:
            from .hubert import HubertEncoder
            if 'state_dict' not in enc_kw:
                path = ck.resolve(cfg.weights, ck.ENV_HUBERT)
                if path:
                    enc_kw['state_dict'] = ck.load_hubert_state_dict(path)
                else:
                    need('mHuBERT checkpoint', ck.ENV_HUBERT)
            if 'codebook' not in enc_kw:
                path = ck.resolve(cfg.quantizer_path, ck.ENV_KMEANS)
                if path:
                    enc_kw['codebook'] = ck.load_kmeans_centroids(path)
                else:
                    need('k-means centres (joblib)', ck.ENV_KMEANS)
            self.encoder = HubertEncoder(config=cfg, device=self.device, **enc_kw)
            return
        if 'state_dict' not in enc_kw:
            path = ck.resolve(getattr(cfg, 'weights', None), ck.ENV_W2VBERT)
            if path:
                enc_kw['state_dict'] = ck.load_w2vbert_state_dict(path)
            else:
                need('w2v-BERT 2.0 checkpoint', ck.ENV_W2VBERT)
        if 'codebook' not in enc_kw:
            path = ck.resolve(getattr(cfg, 'quantizer_path', None), ck.ENV_VQ)
            if path:
                enc_kw['codebook'] = (ck.load_kmeans_centroids(path) if path.endswith(('.bin', '.joblib'))
                                      else ck.load_vq_codebook(path))
            else:
                need('quantizer (codebook) checkpoint', ck.ENV_VQ)
        if self.tokenizer_name == Tokenizers.semantic_s:
            from .encoder import SemanticSEncoder
            self.encoder = SemanticSEncoder(config=cfg, device=self.device, **enc_kw)
        else:
            from .encoder import Wav2VecBertEncoder
            self.encoder = Wav2VecBertEncoder(config=cfg, device=self.device, quantize=True, **enc_kw)

    # reference core.py:120-185
    def encode(self, audio, chunk_size: Optional[int] = None) -> torch.Tensor:
        """ndarray / Tensor [1, L] at model_sample_rate, or a path -> int16 CPU tokens [1, K, T]
        ([K, T_total] when a path is encoded with chunk_size, as in the reference)."""
        self.load_encoder()
        if isinstance(audio, np.ndarray):
            assert audio.ndim == 2, "Audio must be 2D array"
            assert audio.shape[0] == 1, "Audio must mono"
            return self._encode_single(torch.from_numpy(audio))
        if isinstance(audio, torch.Tensor):
            assert audio.ndim == 2, "Audio must be 2D array"
            assert audio.shape[0] == 1, "Audio must mono"
            return self._encode_single(audio)
        if isinstance(audio, (os.PathLike, Path, str)) and not isinstance(audio, bytes):
            if chunk_size is None:
                return self._encode_single(aio.read_audio(audio, self.model_sample_rate))
            # chunks are cut at the file's own rate and resampled one by one, as in batch mode and in the reference
            # (utils.py:82-101), so encode(path, chunk_size) and encode_batch_files agree at chunk boundaries
            parts = [self._encode_single(w)[0] for w in aio.read_audio_chunks(audio, self.model_sample_rate, chunk_size)
                     if w.shape[-1] > 0]
            return torch.cat(parts, dim=-1)
        if isinstance(audio, bytes):
            raise NotImplementedError("Encoding bytes not supported yet")
        raise ValueError(f"Unsupported input type {type(audio)}. Should be one of: ndarray, Tensor, PathLike")

    # reference core.py:187-196
    def _encode_single(self, audio: torch.Tensor) -> torch.Tensor:
        if hasattr(self.encoder, 'encode_single'):            # mHuBERT: the processor transform comes first (core.py:188-189)
            return self.encoder.encode_single(audio).cpu()
        input_batch = audio.to(self.device, torch.float32)
        attention_mask = torch.ones_like(input_batch)
        toks = self.encoder(input_batch, attention_mask)
        return toks.cpu()

    # reference core.py:198-289
    def encode_batch_files(self, batch_size: int, outdir, chunk_size: int = 30, num_workers: int = 12,
                           audio_files: Optional[Sequence] = None, audio_dir=None, **dataloader_kwargs) -> None:
        self.load_encoder()
        assert audio_files or audio_dir, "Either audio_files or audio_dir must be provided"
        assert not (audio_files and audio_dir), "Provide either audio_files or audio_dir, not both"
        outdir = aio.sanitize_path(outdir)
        files = [str(f) for f in audio_files] if audio_files else aio.find_audio_files(str(audio_dir))
        num_workers = max(1, min(num_workers, os.cpu_count() or 1, len(files) or 1))
        sr, rate = self.model_sample_rate, self.model_config.model_token_rate
        self.last_stats = encode_files(self.encoder, files, outdir, sr, rate, chunk_size, batch_size, num_workers,
                                       rel_dir=None if audio_files else str(audio_dir))

    def load_decoder(self, **kwargs):
        """reference core.py:291-313.  Only the acoustic decoder is built."""
        if getattr(self, 'decoder', None) is not None:
            return
        if self.tokenizer_name != Tokenizers.acoustic:
            raise NotImplementedError("semantic token -> audio decoding (GPT-2 + Bark, reference decoder.py:79-) is outside "
                                      "the scope of this build (SURVEY.md section 8f)")
        from . import checkpoints as ck
        from .acoustic import AcousticDecoder
        allkw = {**self.kwargs, **kwargs}
        kw = {k: v for k, v in allkw.items() if k in ('state_dict', 'seed')}
        if 'state_dict' not in kw:
            path = ck.resolve(getattr(self.model_config, 'weights', None), ck.ENV_ENCODEC)
            if path:
                kw['state_dict'] = ck.load_encodec_state_dict(path)
            elif not allkw.get('synthetic_weights', False):
                raise FileNotFoundError(f'AudioToken(acoustic).decode: no EnCodec checkpoint (set config.weights or ${ck.ENV_ENCODEC}, '
                                        'or pass synthetic_weights=True)')
        self.decoder = AcousticDecoder(device=self.device, **kw)

    def decode(self, tokens, **kwargs) -> torch.Tensor:
        """tokens (1, num_codebooks, num_tokens) -> audio fp32 CPU (1, num_samples)  (reference core.py:317-357)."""
        self.load_decoder(**kwargs)
        if isinstance(tokens, np.ndarray):
            tokens = torch.from_numpy(tokens)
        elif isinstance(tokens, (os.PathLike, str)):
            tokens = torch.load(tokens, map_location='cpu')
        elif not isinstance(tokens, torch.Tensor):
            raise ValueError(f"Unsupported input type {type(tokens)}. Should be one of: np.ndarray, torch.Tensor, os.PathLike")
        if tokens.dim() == 2:
            tokens = tokens.unsqueeze(0)
        return self.decoder(tokens.to(dtype=torch.long)).cpu()


def encode_files(encoder, files: Sequence[str], outdir: str, sample_rate: int, token_rate: int, chunk_size: int,
                 batch_size: int, num_workers: int, rel_dir: Optional[str], window_rows: Optional[int] = None,
                 on_error: str = 'log') -> Dict[str, float]:
    """The batched file loop as a bounded-memory stream (reference core.py:259-287: DataLoader workers + prefetch +
    per-batch saves):

      reader threads (host only: file read + RIFF parse, at most 2 * num_workers files in flight)
        -> window of whole files (<= window_rows token rows; default 4 ragged batches)
        -> PCM decode + resample on the device, segment rules, length-bucketed ragged batches
        -> encode (batch k+1 is launched before the tokens of batch k are copied back on a side stream)
        -> a file whose segments are all done goes to the writer threads (one atomic .npy write per file).

    Host memory is bounded by the files in flight plus two windows (the one being collected and the one queued behind it);
    nothing is held until the end of the corpus and every file is handed to the writers as soon as its last segment is back.  `on_error`: 'log' (reference behaviour,
    datasets.py:136-137: log the file and continue) or 'raise'."""
    t0 = time.time()
    pad = int(chunk_size * sample_rate)
    row_budget = min(max(1, batch_size) * encoder.rows_for(pad), getattr(encoder, 'max_rows_per_batch', 1 << 30))
    window_rows = int(window_rows) if window_rows else 4 * row_budget
    device = getattr(encoder, 'device', None)
    on_gpu = device is not None and torch.device(device).type == 'cuda'
    errors: Dict[str, str] = {}
    stats = {'files': 0, 'segments': 0, 'audio_seconds': 0.0, 'windows': 0, 'batches': 0, 'peak_window_bytes': 0,
             't_read_wait': 0.0, 't_prepare': 0.0, 't_launch': 0.0, 't_fetch': 0.0}     # host seconds per stage of the loop

    def fail(path, exc):
        errors[path] = repr(exc)
        logger.error('encode_batch_files: skipping %s: %r', path, exc)
        if on_error == 'raise':
            raise exc

    def read(path):
        try:
            if not path.lower().endswith(AUDIO_EXTS):
                raise NotImplementedError(f'unsupported extension: {path}')
            return aio.read_wav_raw(path)
        except Exception as e:  # noqa: BLE001
            return e

    copy_stream, upload_stream, ingest_stream = _side_streams(device) if on_gpu else (None, None, None)
    writers = ThreadPoolExecutor(max(1, min(4, num_workers)))
    write_futs = []

    def write(path, chunks):
        dst = aio.token_path_flat(path, outdir) if rel_dir is None else aio.token_path_rel(path, outdir, rel_dir)
        aio.save_tokens_atomic(dst, [chunks[k] for k in sorted(chunks)])

    def launch(clips, rws):
        """encode one ragged batch; the per-clip token tensors are gathered into ONE device buffer on the compute
        stream and an event marks the point the side stream may copy it from"""
        toks = encoder.encode_packed(clips, pad, rws)
        if not on_gpu:
            return toks, None, None
        flat = torch.cat([t.reshape(-1) for t in toks])
        return [tuple(t.shape) for t in toks], flat, torch.cuda.current_stream(device).record_event()

    def fetch(pending):
        """tokens of a launched batch -> host arrays; waits only for that batch (side stream + event), so the next
        batch, already launched, keeps the device busy meanwhile"""
        shapes, flat, done, idx = pending
        if flat is None:
            return [t.cpu().numpy() for t in shapes], idx
        host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            host.copy_(flat, non_blocking=True)
            ev = copy_stream.record_event()
        flat.record_stream(copy_stream)
        ev.synchronize()
        out, off, hv = [], 0, host.numpy()
        for shp in shapes:
            n = int(np.prod(shp))
            out.append(hv[off:off + n].reshape(shp))
            off += n
        return out, idx

    class WindowState:
        """segments of one window + the bookkeeping that tells when a file is complete"""

        def __init__(self, segs, rows):
            self.segs, self.rows = segs, rows
            self.per_file: Dict[str, Dict[int, np.ndarray]] = {}
            self.left: Dict[str, int] = {}
            for s_ in segs:
                self.left[s_.file_name] = self.left.get(s_.file_name, 0) + 1

        def absorb(self, host, idx):
            for i, t in zip(idx, host):
                s_ = self.segs[i]
                self.per_file.setdefault(s_.file_name, {})[s_.chunk_index] = np.array(t[:, :s_.config.length_tokens])
                stats['audio_seconds'] += s_.config.length_seconds
                self.left[s_.file_name] -= 1
                if self.left[s_.file_name] == 0:               # every chunk of the file is encoded: write it now
                    write_futs.append((s_.file_name, writers.submit(write, s_.file_name, self.per_file.pop(s_.file_name))))
                    stats['files'] += 1

    def prepare(window, raws=None, out_alloc=None):
        """window: list of (path, sr, pcm) -> WindowState (PCM decode + resampling on the device, segment rules);
        raws[i] = the payload of file i already on the device (prepare_async) or None; out_alloc = bump allocator over
        the window's waveform buffer."""
        segs, nbytes = [], 0
        for wi, (path, sr, pcm) in enumerate(window):
            try:
                nbytes += pcm.nbytes
                for ci, wave in enumerate(aio.convert_chunks(sr, pcm, sample_rate, chunk_size, device if on_gpu else None,
                                                             raw=raws[wi] if raws else None,
                                                             out_alloc=out_alloc if raws and raws[wi] is not None else None)):
                    for s_ in aio.iter_segments(wave, path, sample_rate, token_rate, chunk_size):
                        s_.chunk_index = ci                # one segment per streamed chunk (datasets.py:88-105)
                        segs.append(s_)
            except Exception as e:  # noqa: BLE001
                segs = [s_ for s_ in segs if s_.file_name != path]
                fail(path, e)
        stats['peak_window_bytes'] = max(stats['peak_window_bytes'], nbytes)
        rows = [encoder.rows_for_tokens(length_tokens(int(s_.wave.numel()), sample_rate, token_rate), pad) for s_ in segs]
        stats['segments'] += len(segs)
        stats['windows'] += 1
        return WindowState(segs, rows)

    def prepare_async(window):
        """prepare() on side streams: the PCM upload and the decode / resampling kernels of the NEXT window run while
        the compute stream is busy with the batches of the current one"""
        if not on_gpu:
            return prepare(window), None
        # ONE device buffer per window for the PCM payloads and one for the decoded waveforms: per-file tensors would be
        # per-file cudaMallocs whenever the side streams' allocator pools are cold (measured: 2.3 s of 3.3 s for 1250 files)
        ok = [pcm.dtype in (np.int16, np.float32) and pcm.ndim <= 2 for _, _, pcm in window]
        offs, total = [], 0
        for (_, _, pcm), good in zip(window, ok):
            offs.append(total)
            if good:
                total += (pcm.nbytes + 15) // 16 * 16
        n_out = sum(sum((n + 3) // 4 * 4 for n in aio.chunk_output_lengths(sr, int(pcm.shape[0]), sample_rate, chunk_size))
                    for (_, sr, pcm), good in zip(window, ok) if good)
        with torch.cuda.stream(upload_stream):
            raw_all = torch.empty(max(total, 16), dtype=torch.uint8, device=device)
            raws = []
            for (_, _, pcm), good, off in zip(window, ok, offs):
                try:
                    raws.append(aio.upload_pcm(pcm, device, raw_all[off:off + pcm.nbytes]) if good else None)
                except Exception:  # noqa: BLE001 -- prepare() meets the same error and reports it for the file
                    raws.append(None)
            uploaded = upload_stream.record_event()
        with torch.cuda.stream(ingest_stream):
            ingest_stream.wait_event(uploaded)
            raw_all.record_stream(ingest_stream)
            wave_all = torch.empty(max(n_out, 1), dtype=torch.float32, device=device)
            cursor = [0]

            def out_alloc(n):
                a = cursor[0]
                cursor[0] = a + (n + 3) // 4 * 4                    # every chunk starts 16-byte aligned
                return wave_all[a:a + n].view(1, n)

            st_ = prepare(window, raws, out_alloc)
            ready = ingest_stream.record_event()
        return st_, ready

    def launch_window(st_, ready):
        """all ragged batches of a prepared window, back to back (no host synchronisation in between)"""
        comp = torch.cuda.current_stream(device) if on_gpu else None
        if ready is not None:
            comp.wait_event(ready)
            for s_ in st_.segs:
                if s_.wave.is_cuda:                      # 8-bit / 32-bit PCM is decoded on the host (io.convert_chunks)
                    s_.wave.record_stream(comp)
        out = []
        for idx in bucket_by_rows(st_.rows, row_budget):
            out.append((*launch([st_.segs[i].wave for i in idx], [st_.rows[i] for i in idx]), idx))
            stats['batches'] += 1
        return out

    def est_rows(sr, pcm):
        return length_tokens(int(pcm.shape[0] * sample_rate / max(sr, 1)), sample_rate, token_rate)

    max_in_flight = max(2, 2 * num_workers)
    try:
        with ThreadPoolExecutor(num_workers) as readers:
            it = iter(files)
            in_flight = []                       # (path, future) in file order

            def top_up():
                while len(in_flight) < max_in_flight:
                    path = next(it, None)
                    if path is None:
                        return
                    in_flight.append((path, readers.submit(read, path)))

            carry = []

            def next_window(limit=None):
                """whole files up to `limit` (default window_rows) token rows (None at the end); the readers keep prefetching"""
                limit = window_rows if limit is None else limit
                window, w_rows = list(carry), sum(est_rows(sr, pcm) for _, sr, pcm in carry)
                carry.clear()
                top_up()
                while in_flight:
                    path, fut = in_flight.pop(0)
                    res = fut.result()
                    top_up()
                    if isinstance(res, Exception):
                        fail(path, res)
                        continue
                    sr, pcm = res
                    r = est_rows(sr, pcm)
                    if window and w_rows + r > limit:
                        carry.append((path, sr, pcm))
                        return window
                    window.append((path, sr, pcm))
                    w_rows += r
                return window or None

            # software pipeline over windows: launch every batch of window w, prepare window w + 1 (host work + side
            # streams) while the device encodes AND queue its batches behind those of window w, only then collect the tokens
            # of window w and hand finished files to the writers — the device never waits for the host at a window boundary
            # (planning + ~700 launches of the first batch of a window are 10-20 ms of host time)
            def timed(key, fn, *a):
                t = time.perf_counter()
                r = fn(*a)
                stats[key] += time.perf_counter() - t
                return r

            # the first window is one batch: the device starts after a quarter of the read + decode latency of a full window
            w0 = timed('t_read_wait', next_window, min(window_rows, row_budget))
            cur = timed('t_prepare', prepare_async, w0) if w0 else None
            launched = timed('t_launch', launch_window, *cur) if cur is not None else None
            while cur is not None:
                nxt_files = timed('t_read_wait', next_window)
                nxt = timed('t_prepare', prepare_async, nxt_files) if nxt_files else None
                nxt_launched = timed('t_launch', launch_window, *nxt) if nxt is not None else None
                for item in launched:
                    cur[0].absorb(*timed('t_fetch', fetch, item))
                cur, launched = nxt, nxt_launched
    finally:
        writers.shutdown(wait=True)
    for path, f in write_futs:
        e = f.exception()
        if e is not None:
            stats['files'] -= 1
            fail(path, e)
    wall = time.time() - t0
    return {**stats, 'wall_seconds': wall, 'audio_seconds_per_second': stats['audio_seconds'] / wall if wall > 0 else 0.0,
            'errors': errors}
