"""Batch sharding for multi-GPU runs: one process per GPU, instances partitioned contiguously, the operator
replicated on every GPU, no data-path collective.  The only exchange is a sum of a few scalars
(global energy / error norms) -- `Context.allreduce_sum` (NCCL) or `torch.distributed.all_reduce`."""
from __future__ import annotations


def shard_range(B: int, world_size: int, rank: int, multiple: int = 1) -> tuple[int, int]:
    """Instances [b0, b1) owned by `rank`: contiguous, balanced to within one block of `multiple` instances
    (use multiple = G for periodic operator tables so that b % G is preserved on every shard)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    blocks = (B + multiple - 1) // multiple
    base, rem = divmod(blocks, world_size)
    start_blk = rank * base + min(rank, rem)
    nblk = base + (1 if rank < rem else 0)
    b0 = min(B, start_blk * multiple)
    b1 = min(B, (start_blk + nblk) * multiple)
    return b0, b1
