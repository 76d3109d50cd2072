"""Multi-GPU plumbing shared by bench.py and the tests: how reads are dealt to ranks and how
the per-rank counter vectors are merged.

The path shards by reads: every rank holds the whole pattern table and a private
uint32[2*n_patterns] counter vector; the only exchange is one all-reduce (sum) of those vectors
(SURVEY 8e).  uint32 addition wraps, and so does the int32 addition NCCL/gloo perform on the
same bits, so the merged result is bit-identical to a single-GPU run."""
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
