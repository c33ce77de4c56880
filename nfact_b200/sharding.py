"""Row sharding of the connectivity matrix across GPUs (one process per GPU).

X is partitioned by seed rows; rank r owns rows [row_range(m, r, P)).  The W shard stays local, H is
replicated.  Per NMF iteration only the k x n numerator W'X, the k x k Gram W'W and one float64 scalar
cross ranks (all-reduce, issued by libnfact_b200 on the compute stream once ``Handle.init_nccl`` ran).
"""
from __future__ import annotations


def row_range(m: int, rank: int, world: int):
    """Contiguous, balanced row range of ``rank``: sizes differ by at most one row."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    return (m * rank) // world, (m * (rank + 1)) // world


def init_nccl_from_torch(handle, group=None) -> None:
    """Create the library's NCCL communicator, using torch.distributed only to broadcast the unique id."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def bcast(payload):
        obj = [payload]
        dist.broadcast_object_list(obj, src=0, group=group)
        return obj[0]
    handle.init_nccl(rank, world, bcast)
    handle.group = group      # the host-side collectives of the sharded initialisation use the same world
