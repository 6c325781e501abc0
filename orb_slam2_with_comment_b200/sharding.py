"""Frame / frame-pair sharding across the GPUs of one box.

The path has no exchange step: rank g owns the contiguous range [g*n/G, (g+1)*n/G) of the global work list, runs the
same single-GPU calls on it, and the host concatenates per-rank results in rank order — which is the input order, so
no permutation is ever needed.  Nothing here touches a GPU; `gather_ragged` is the only collective and moves results,
not data-path operands.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one), order-preserving."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_ragged(local: np.ndarray, dist=None) -> np.ndarray:
    """Concatenate every rank's (ragged, first-axis) array in rank order on every rank.  `dist` = torch.distributed
    (or None / uninitialised for a single process)."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return np.concatenate(parts, axis=0)


def merge_frame_results(counts_per_rank, kp_per_rank, desc_per_rank):
    """Per-rank (counts [b_r], kp [b_r, cap], desc [b_r, cap, 32]) -> global arrays in input order plus the exclusive
    scan of the counts (offset of every frame's keypoints in a packed list)."""
    counts = np.concatenate(counts_per_rank)
    kp = np.concatenate(kp_per_rank)
    desc = np.concatenate(desc_per_rank)
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return counts, kp, desc, off
