"""Multi-GPU plumbing shared by bench.py and the tests: how reads are dealt to ranks and how
the per-rank counter vectors are merged.

The path shards by reads: every rank holds the whole pattern table; the only exchange is the
sum of the uint32[2*n_patterns] counter vectors.  On NVLink-connected GPUs there is
no exchange step at all: rank 0 exports its vector (CUDA IPC), the others attach it and their
kernels add their (rare) hits straight into it (share_counters).  Where that is not possible
the private vectors are summed with one all-reduce (all_reduce_counts).  uint32 addition wraps,
and so does the int32 addition NCCL/gloo perform on the same bits, so either way the merged
result is bit-identical to a single-GPU run."""
from __future__ import annotations

from typing import List, Sequence


def deal_round_robin(n_items: int, rank: int, world: int) -> range:
    """Items (blocks of reads) rank `rank` of `world` processes: i = rank, rank + world, ..."""
    return range(rank, n_items, world)


def split_reads(reads: Sequence, rank: int, world: int, block: int = 1024) -> List:
    """Blocks of `block` consecutive reads dealt round-robin, like the CLI deals staging blocks
    to devices."""
    out = []
    n_blocks = (len(reads) + block - 1) // block
    for b in deal_round_robin(n_blocks, rank, world):
        out.extend(reads[b * block:(b + 1) * block])
    return out


def all_reduce_counts(counts, dist=None):
    """Sum a torch int32 tensor holding uint32 counter bits over all ranks, in place."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts)
    return counts


def strong_share(total: int, rank: int, world: int) -> int:
    """Reads of a fixed total that rank `rank` scans (strong scaling): equal shares, the first
    ranks take the remainder."""
    return total // world + (1 if rank < total % world else 0)


def share_counters(engine, dist=None, rank: int = 0) -> str:
    """One counter vector for all ranks: rank 0's, attached by the others over CUDA IPC.  Returns
    "peer-atomics" when every rank attached, "all_reduce" when any could not (each rank then keeps
    its own vector and the caller merges with all_reduce_counts)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return "single"
    box = [engine.export_counters() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ok = 1
    if rank != 0:
        try:
            engine.attach_counters(box[0])
        except Exception:
            ok = 0
    flags = [None] * dist.get_world_size()
    dist.all_gather_object(flags, ok)
    if all(flags):
        return "peer-atomics"
    if rank != 0 and ok:
        raise RuntimeError("some ranks attached rank 0's counters and some could not")
    return "all_reduce"
