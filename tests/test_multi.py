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


# ---------------------------------------------------------------------------------------------
# counting mode: owners partition the hash space, the per-owner histograms add up


def _kc_rank(rank, world, port, k, reads, q):
    """what one rank of tools/kc_bench.py does, with the oracle standing in for the device: extract
    the k-mers of its share of the reads, keep those it owns and the lists for the others, exchange
    the lists (the staged form's all-to-all, spelt as an all-gather because gloo lacks it), count what it owns, all-reduce the 256 bins"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kco = util.KcOracle()
    mine = sharding.split_reads(reads, rank, world, block=64)
    hashed = np.concatenate([kco.hashed_kmers(r, k) for r in mine if len(r) >= k] or [np.zeros(0, np.uint64)])
    own = util.kcgpu.owner_of(hashed, world)
    send = [hashed[own == p] for p in range(world)]
    everybody = [None] * world
    dist.all_gather_object(everybody, send)  # gloo has no all_to_all: everybody sees every list and takes its own
    owned = np.concatenate([everybody[src][rank] for src in range(world)])
    assert np.all(util.kcgpu.owner_of(owned, world) == rank)
    hist, n_inst, n_dist = kco.count_hashed(owned, k)
    t = torch.from_numpy(hist.view(np.int64).copy())
    extra = torch.tensor([n_inst, n_dist], dtype=torch.int64)
    dist.all_reduce(t)
    dist.all_reduce(extra)
    if rank == 0:
        q.put((t.numpy().view(np.uint64).copy(), extra.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_counting_mode_two_owners_add_up():
    k = 31
    rng = np.random.default_rng(6)
    reads = util.make_genome_reads(rng, 15000, 1200, jitter=30, repeat=4)
    want, n_inst, n_dist = util.KcOracle().count_reads(reads, k)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_kc_rank, args=(r, 2, port, k, reads, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, (inst, dist_) = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(merged, want) and inst == n_inst and dist_ == n_dist
