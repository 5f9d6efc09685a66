"""Batch sharding of the CSPN path over the GPUs of one box.

Every image is independent (no op in ``CSPN_new.py:26-128`` / ``CSPN_ours.py:24-54`` mixes the batch
dimension), so the path shards as contiguous batch slices - exactly what the reference's DataParallel
scatter does (``network/libs/base/encoding.py:134``) - with NO collective on the CSPN path.
"""
from __future__ import annotations


def batch_slice(global_batch: int, world_size: int, rank: int) -> slice:
    """Contiguous, balanced slice of ``range(global_batch)`` owned by ``rank`` (first ranks get the remainder)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard(tensor, world_size: int, rank: int):
    """The rank's batch slice of ``tensor`` (``None`` passes through)."""
    if tensor is None:
        return None
    return tensor[batch_slice(tensor.shape[0], world_size, rank)]
