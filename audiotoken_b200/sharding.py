"""Collective-free data parallelism: the file list is split across ranks, one process per GPU.

Every (file, chunk) is encoded independently (reference datasets.py:75-105: fresh front-end statistics
per segment, no cross-clip state), so there is no exchange step on the hot path and no collective is
needed.  Files are assigned by greedy longest-processing-time on their duration so that
all ranks finish together; output files are disjoint per rank.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence


def lpt_shards(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Indices per rank; deterministic (ties broken by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [sorted(s) for s in shards]


def shard_files(files: Sequence[str], durations: Sequence[float], world_size: int, rank: int) -> List[str]:
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} not in [0, {world_size})')
    return [files[i] for i in lpt_shards(durations, world_size)[rank]]
