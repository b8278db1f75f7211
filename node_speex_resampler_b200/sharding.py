"""Stream sharding across the GPUs of one box (BASELINE.json configs[4]).

Streams are independent (no cross-stream term anywhere in deps/speex/resample.c: the channel
loop at :1070-1078 touches only that channel's state), so the multi-GPU path is a partition of
the stream list: rank r of `world` owns one contiguous block, keeps its own filter-bank
replica, state arrays, CUDA streams and pinned staging, and there is no collective on the
data path. torch.distributed is used by bench.py only for the start barrier and the
max-over-ranks of the timed region."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the streams rank `rank` owns: contiguous blocks, sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_streams: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(n_streams, world, r) for r in range(world))]


def owner_of(stream: int, n_streams: int, world: int) -> int:
    """rank that owns global stream index `stream`"""
    for r in range(world):
        lo, hi = shard_range(n_streams, world, r)
        if lo <= stream < hi:
            return r
    raise ValueError("stream out of range")


def split_chunks(chunks: Sequence, world: int, rank: int) -> Sequence:
    """this rank's slice of a per-stream list (chunks, resamplers, results ...)"""
    lo, hi = shard_range(len(chunks), world, rank)
    return chunks[lo:hi]
