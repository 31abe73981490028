"""Column sharding of a batch over GPUs / ranks — the only multi-GPU logic the path needs.

Columns (configurations) are independent (reference: the `schedule(static)` loop of
algorithm/parallel/rnea.hpp:74-82 has no synchronisation), so a batch splits into contiguous column ranges,
one per device, with no collective.  `column_range` is the partition the C ABI applies inside a multi-device
pool (capi.cu run_call) and the one a one-process-per-GPU launcher (torchrun) applies across ranks.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import numpy as np


def column_range(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[c0, c1) of `rank`: ceil(batch / world_size) columns each, the tail ranks may be short or empty."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("rank must be in [0, world_size)")
    per = -(-batch // world_size)
    c0 = min(batch, rank * per)
    return c0, min(batch, c0 + per)


def shard_columns(arrays: Sequence[np.ndarray], world_size: int, rank: int):
    """Views of this rank's columns of every (rows x B) block."""
    B = arrays[0].shape[1]
    for a in arrays:
        if a.shape[1] != B:
            raise ValueError("wrong argument size: all blocks must have the same number of columns")
    c0, c1 = column_range(B, world_size, rank)
    return [a[:, c0:c1] for a in arrays]


def run_sharded(fn: Callable[..., np.ndarray], arrays: Sequence[np.ndarray], world_size: int, rank: int) -> np.ndarray:
    """Apply a batched evaluation `fn(*blocks) -> (rows x b)` to this rank's shard."""
    return fn(*shard_columns(arrays, world_size, rank))
