"""Multi-GPU driver: one process per GPU, the file list sharded by duration, no collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m audiotoken_b200.launch --tokenizer semantic_m --indir corpus/ --outdir tokens/

Each rank reads the RIFF headers of all files (cheap), computes the same longest-processing-time
assignment, and encodes only its own shard into the shared output directory.  `torch.distributed`
(NCCL on GPUs, gloo in CPU tests) is used only for the start/end barrier and for summing the per-rank
statistics; the encode path itself has no exchange step.
"""
from __future__ import annotations

import argparse
import json
import os

import torch


def run(args) -> dict:
    from . import io as aio
    from .core import AudioToken
    from .sharding import shard_files

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    files = aio.find_audio_files(args.indir)
    durations = []
    for f in files:
        try:
            sr, n, _ = aio.wav_info(f)
            durations.append(n / sr)
        except Exception:  # noqa: BLE001
            durations.append(0.0)
    mine = shard_files(files, durations, world, rank)
    tok = AudioToken(tokenizer=args.tokenizer, device=f'cuda:{local}', synthetic_weights=args.synthetic_weights)
    stats = {'files': 0, 'audio_seconds': 0.0, 'wall_seconds': 0.0}
    if mine:
        # audio_dir semantics (relative output layout) on an explicit shard of the directory
        from .core import encode_files
        tok.load_encoder()
        stats = encode_files(tok.encoder, mine, aio.sanitize_path(args.outdir), tok.model_sample_rate,
                             tok.model_config.model_token_rate, args.chunk_size, args.batch_size, args.num_workers,
                             rel_dir=str(args.indir))
    return {'rank': rank, 'world': world, **{k: stats[k] for k in ('files', 'audio_seconds', 'wall_seconds')}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tokenizer', default='semantic_m')
    ap.add_argument('--indir', required=True)
    ap.add_argument('--outdir', required=True)
    ap.add_argument('--batch_size', type=int, default=64)
    ap.add_argument('--chunk_size', type=int, default=30)
    ap.add_argument('--num_workers', type=int, default=12)
    ap.add_argument('--synthetic-weights', action='store_true',
                    help='seeded synthetic weights instead of checkpoints (benchmarks; see audiotoken_b200/checkpoints.py)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    res = run(args)
    if world > 1:
        import torch.distributed as dist
        out = [None] * world
        dist.all_gather_object(out, res)
        if res['rank'] == 0:
            total = sum(r['audio_seconds'] for r in out)
            wall = max(r['wall_seconds'] for r in out)
            print(json.dumps({'ranks': out, 'audio_seconds': total, 'wall_seconds': wall,
                              'audio_seconds_per_second': total / wall if wall else 0.0}))
        dist.destroy_process_group()
    else:
        print(json.dumps(res))


if __name__ == '__main__':
    main()
