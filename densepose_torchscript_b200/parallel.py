"""Multi-GPU sharding of independent images / video frames.

The forward pass of one image never talks to another image, so scaling out is pure data sharding: one
process per GPU (torch.distributed for the plumbing), each with its own weight replica, streams and
workspace; frames are dealt to ranks, every rank runs them through its own engine, and the per-frame
results are gathered HOST-side in submission order.  There is no collective on the data path — NCCL /
NVLink are only touched by the optional barrier / timing reductions in bench.py.
"""
from typing import Any, Callable, List, Optional, Sequence

import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int, block: int = 1) -> List[int]:
    """Indices owned by `rank`: blocks of `block` consecutive frames dealt round-robin (block = the per-GPU
    batch keeps batches contiguous in time for video)."""
    if world <= 0 or not (0 <= rank < world) or block <= 0:
        raise ValueError("bad shard arguments")
    out = []
    for start in range(rank * block, n_items, world * block):
        out.extend(range(start, min(start + block, n_items)))
    return out


def run_sharded(items: Sequence[Any], process: Callable[[List[Any]], List[Any]], batch: int = 8,
                group: Optional[Any] = None, dst: int = 0) -> Optional[List[Any]]:
    """Runs `process` over this rank's share of `items` in batches of `batch`; returns the full result list in
    the original order on rank `dst` (None elsewhere). Results must be picklable host objects."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard_indices(len(items), rank, world, batch)
    local = []
    for i in range(0, len(mine), batch):
        idx = mine[i:i + batch]
        res = process([items[j] for j in idx])
        if len(res) != len(idx):
            raise RuntimeError("process() must return one result per item")
        local.extend(zip(idx, res))
    if world == 1:
        merged = local
    else:
        gathered = [None] * world if rank == dst else None
        dist.gather_object(local, gathered, dst=dst, group=group)      # host-side gather, not on the data path
        if rank != dst:
            return None
        merged = [p for part in gathered for p in part]
    merged.sort(key=lambda p: p[0])
    if [p[0] for p in merged] != list(range(len(items))):
        raise RuntimeError("sharding lost or duplicated frames")
    return [p[1] for p in merged]
