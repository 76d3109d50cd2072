"""The N > 1 path on CPU: two gloo ranks shard the reads, count their shards (with the
test-only host emulation of the kernel, since there is no GPU here) and merge the counter
vectors with the same all-reduce bench.py issues over NCCL.  The merged result must equal the
oracle's count of all the reads, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from util import vafgpu

sys.path.insert(0, util.PKG)
import sharding  # noqa: E402


def _rank(rank, world, port, pf, k, n_pat, reads, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    mine = sharding.split_reads(reads, rank, world, block=64)
    got, _, _ = util.Sim().count(k, keys, vals, n_pat, util.pack_stream(mine, k))
    t = torch.from_numpy(got.view(np.int32).copy())
    t[0] += -5 if rank == 0 else 5                       # prove the reduction really runs
    t[1] = np.int32(-(2 ** 31)) if rank == 0 else t[1]   # and that wrap-around is harmless
    sharding.all_reduce_counts(t, dist)
    if rank == 0:
        q.put((t.numpy().view(np.uint32).copy(), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_merge_to_the_single_rank_result(tmp_path, oracle):
    k, n_pat = 21, 200
    rng = np.random.default_rng(4)
    pats = util.make_patterns(rng, n_pat, k)
    reads = util.make_reads(rng, pats, k, 3000, plant=0.8, jitter=20)
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    want, _, _ = oracle.count_reads(pf, k, reads)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_rank, args=(r, 2, port, pf, k, n_pat, reads, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert 0 < n0 < len(reads)
    expect = want.copy()
    # rank 1 counted want[1] - (rank 0's share); rank 0's word 1 was replaced by 2^31
    shard0, _, _ = oracle.count_reads(pf, k, sharding.split_reads(reads, 0, 2, block=64))
    expect[1] = np.uint32((int(want[1]) - int(shard0[1]) + 2 ** 31) % 2 ** 32)
    assert np.array_equal(merged, expect)


def test_dealing_covers_every_read_once():
    reads = list(range(10007))
    for world in (1, 2, 3, 8):
        got = sorted(sum((sharding.split_reads(reads, r, world, block=100) for r in range(world)), []))
        assert got == reads
