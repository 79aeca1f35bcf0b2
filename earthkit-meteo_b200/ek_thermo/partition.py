"""Field partitioner: shard the flat grid-point range across the GPUs of one box.

Every point is independent, so there is no exchange step and no collective:
rank r of world W owns the contiguous range ``shard_range(n, W, r, align)`` of every input and
writes the same range of every output; results stay resident on their GPU.  ``align`` keeps shard
edges on slab boundaries (e.g. one member x level slab of an ENS field) and 16-byte aligned.
"""
from __future__ import annotations

import os

from . import _backend as _b


def shard_range(n: int, world: int, rank: int, align: int = 1):
    """[begin, end) of `rank`'s shard (native: ek_thermo_shard_range)."""
    return _b.shard_range(int(n), int(world), int(rank), int(align))


def all_shards(n: int, world: int, align: int = 1):
    return [shard_range(n, world, r, align) for r in range(world)]


def env_rank_world():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_view(x, world: int, rank: int, align: int = 1):
    """The slice of a flat (1-D) tensor owned by `rank`."""
    if x.dim() != 1:
        raise ValueError("shard_view expects a flat tensor; use x.reshape(-1)")
    b, e = shard_range(x.numel(), world, rank, align)
    return x[b:e]
