"""Parity of the counting mode's CUDA path (through the C ABI of include/kcgpu.h) with the
oracle and with the reference's golden histograms.  Needs a B200: -m gpu.

Integer work: every comparison is bit-exact."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from util import kcgpu

pytestmark = pytest.mark.gpu

KC_CLI = os.path.join(util.PKG, "kc-c4")
GOLDEN_KC = os.path.join(util.GOLDEN, "kc")


@pytest.fixture(scope="module")
def kco():
    return util.KcOracle()


@pytest.fixture(scope="module")
def torch_cuda(lib):
    import torch
    assert torch.cuda.is_available()
    return torch


def reads_case(seed, **kw):
    rng = np.random.default_rng(seed)
    reads = util.make_genome_reads(rng, kw.pop("genome", 40000), kw.pop("n", 5000), **kw)
    return reads + [b"", b"ACGT", b"N" * 64, b"ACGTTGCATTGACCA" * 4]


@pytest.mark.parametrize("k", [1, 2, 5, 11, 15, 16, 21, 27, 28, 31])
def test_add_read_matches_oracle_over_k(kco, lib, k):
    reads = reads_case(200 + k, jitter=100, junk_rate=0.004, lower_rate=0.03, n_rate=0.01, repeat=15)
    want, n_inst, n_dist = kco.count_reads(reads, k)
    with kcgpu.Counter(k, 1 << 21, block_bytes=1 << 17) as c:
        for r in reads:
            c.add_read(r)
        got, st = c.histogram()
    assert np.array_equal(got, want)
    assert st["n_kmers"] == n_inst and st["n_distinct"] == n_dist and st["n_overflow"] == 0
    assert st["n_reads"] == sum(len(r) >= k for r in reads)
    assert st["n_bases"] == sum(len(r) for r in reads if len(r) >= k)
    assert st["n_blocks"] > 3


@pytest.mark.parametrize("list_slots", [kcgpu.NO_LISTS, 1 << 10, 1 << 16])
def test_list_sizes_do_not_change_the_result(kco, lib, list_slots):
    """no lists at all, lists that flush every few kilobytes, lists that flush a few times"""
    for k in (9, 31):
        reads = reads_case(300 + k, jitter=50, repeat=8)
        want, n_inst, n_dist = kco.count_reads(reads, k)
        with kcgpu.Counter(k, 1 << 21, block_bytes=1 << 17, list_slots=list_slots) as c:
            for r in reads:
                c.add_read(r)
            got, st = c.histogram()
        assert np.array_equal(got, want) and st["n_kmers"] == n_inst and st["n_distinct"] == n_dist
        if list_slots == kcgpu.NO_LISTS:
            assert st["n_flushes"] == 0 and st["list_slots"] == 0
        else:
            assert st["n_flushes"] >= (2 if list_slots == 1 << 16 else 5)


def test_full_lists_fall_back_to_the_table(kco, torch_cuda):
    """a context whose owners the caller named never flushes by itself: what does not fit its
    lists goes straight to the table, nothing is lost"""
    torch = torch_cuda
    k = 31
    reads = reads_case(17, repeat=4)
    d = torch.from_numpy(util.pack_stream_strict(reads, k)).cuda()
    want, n_inst, n_dist = kco.count_reads(reads, k)
    with kcgpu.Counter(k, 1 << 21, list_slots=1 << 15) as c:
        c.set_owners(0, [None])
        c.count_device(d.data_ptr(), d.numel())
        got, st = c.histogram()
    assert np.array_equal(got, want) and st["n_kmers"] == n_inst and st["n_distinct"] == n_dist
    assert st["n_direct"] > n_inst // 2
    assert st["n_flushes"] <= 2  # the one set_owners does, the one histogram does: none in between


def test_saturating_count(kco, lib):
    """one k-mer thousands of times, from every lane at once: 1023 is the ceiling (kc-c4.c:125)"""
    reads = [b"A" * 5000, b"T" * 3000, b"ACGT" * 600]
    for k in (4, 11, 31):
        want, n_inst, n_dist = kco.count_reads(reads, k)
        with kcgpu.Counter(k, 1 << 12) as c:
            for r in reads:
                c.add_read(r)
            got, st = c.histogram()
        assert np.array_equal(got, want) and st["n_kmers"] == n_inst and st["n_distinct"] == n_dist
        assert got[255] >= 1


def test_long_reads_are_cut_with_overlap(kco, lib):
    rng = np.random.default_rng(3)
    reads = util.make_genome_reads(rng, 300000, 3, mean_len=250000, n_rate=0.0005)
    for k in (13, 31):
        want, n_inst, _ = kco.count_reads(reads, k)
        with kcgpu.Counter(k, 1 << 21, block_bytes=1 << 16) as c:
            for r in reads:
                c.add_read(r)
            got, st = c.histogram()
        assert np.array_equal(got, want) and st["n_kmers"] == n_inst


def test_device_stream_and_reset(kco, torch_cuda):
    torch = torch_cuda
    reads = reads_case(11, lower_rate=0.1)
    k = 21
    stream = util.pack_stream_strict(reads, k)
    d = torch.from_numpy(stream).cuda()
    want, n_inst, n_dist = kco.count_reads(reads, k)
    with kcgpu.Counter(k, 1 << 21) as c:
        c.count_device(d.data_ptr(), d.numel())
        got, st = c.histogram()
        assert np.array_equal(got, want) and st["n_kmers"] == n_inst
        # counting the same stream again doubles every count: hist2[2c] = hist1[c]
        c.count_device(d.data_ptr(), d.numel())
        got2, _ = c.histogram()
        want2 = np.zeros(256, dtype=np.uint64)
        for cnt in range(1, 256):
            want2[min(2 * cnt, 255)] += want[cnt] if cnt < 255 else 0
        want2[255] += want[255]
        assert np.array_equal(got2, want2)
        c.reset()
        empty, st0 = c.histogram()
        assert int(empty.sum()) == 0 and st0["n_kmers"] == 0
        c.count_device(d.data_ptr(), d.numel())
        again, _ = c.histogram()
        assert np.array_equal(again, want)
    with pytest.raises(util.vafgpu.VafGpuError):
        with kcgpu.Counter(k, 1 << 12) as c:
            c.count_device(d.data_ptr() + 1, d.numel() - 16)


def test_host_stream_pinned_and_pageable(kco, torch_cuda):
    """kcgpu_submit_stream: the caller's packed stream from page-locked memory (copied as it is)
    and from pageable memory (through the staging blocks), cut into blocks at read boundaries"""
    torch = torch_cuda
    rng = np.random.default_rng(21)
    reads = util.make_genome_reads(rng, 60000, 4000, jitter=70, lower_rate=0.05, n_rate=0.01)
    reads += util.make_genome_reads(rng, 300000, 2, mean_len=200000)  # longer than a block: cut with overlap
    for k in (17, 31):
        want, n_inst, n_dist = kco.count_reads(reads, k)
        buf = b"".join(r + b"\n" for r in reads if len(r) >= k)[:-1]  # no trailing separator: the engine adds it
        for pinned in (True, False):
            t = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
            if pinned:
                t = t.pin_memory()
            with kcgpu.Counter(k, 1 << 21, block_bytes=1 << 16) as c:
                c.submit_stream(t.data_ptr(), t.numel())
                got, st = c.histogram()
            assert np.array_equal(got, want) and st["n_kmers"] == n_inst and st["n_distinct"] == n_dist, (k, pinned)
            assert st["n_blocks"] > 10


def test_full_table_is_reported_not_walked_forever(torch_cuda):
    torch = torch_cuda
    reads = reads_case(12, genome=200000, n=3000)
    stream = util.pack_stream_strict(reads, 31)
    d = torch.from_numpy(stream).cuda()
    with kcgpu.Counter(31, 4096) as c:
        c.count_device(d.data_ptr(), d.numel())
        hist, st = c.histogram()
    # (4096 slots asked for; k = 31 needs 2^9 regions of at least 16 slots for the tag to fit: 8192)
    assert st["table_slots"] == 8192
    assert st["n_overflow"] > 0 and st["n_distinct"] <= st["table_slots"] and int(hist.sum()) == st["n_distinct"]


@pytest.mark.parametrize("n_parts", [1, 2, 3, 8])
def test_extract_exchange_insert(kco, torch_cuda, n_parts):
    """the staged several-GPU form on one device: per-owner lists hold exactly the hashed k-mers
    the oracle extracts, and inserting every list into its owner's table gives the histogram"""
    torch = torch_cuda
    k = 31
    reads = reads_case(13, repeat=4)
    stream = util.pack_stream_strict(reads, k)
    d = torch.from_numpy(stream).cuda()
    hashed = np.concatenate([kco.hashed_kmers(r, k) for r in reads if len(r) >= k])
    want, n_inst, n_dist = kco.count_reads(reads, k)
    cap = hashed.size
    keys = torch.zeros(n_parts * cap, dtype=torch.int64, device="cuda")
    counts = torch.zeros(n_parts, dtype=torch.int32, device="cuda")
    owners = [kcgpu.Counter(k, 1 << 20) for _ in range(n_parts)]
    try:
        owners[0].extract_device(d.data_ptr(), d.numel(), n_parts, keys.data_ptr(), cap, counts.data_ptr())
        owners[0].sync()
        got_counts = counts.cpu().numpy()
        own = kcgpu.owner_of(hashed, n_parts)
        total = np.zeros(256, dtype=np.uint64)
        dist = 0
        for p in range(n_parts):
            part = keys[p * cap: p * cap + int(got_counts[p])]
            assert np.array_equal(np.sort(part.cpu().numpy().view(np.uint64)), np.sort(hashed[own == p]))
            owners[p].insert_device(part.data_ptr(), part.numel(), n_parts)
            h, st = owners[p].histogram()
            total += h
            dist += st["n_distinct"]
        assert np.array_equal(total, want) and dist == n_dist
        # lists too short: the excess is dropped and reported, never written out of bounds
        counts.zero_()
        owners[0].extract_device(d.data_ptr(), d.numel(), n_parts, keys.data_ptr(), 100, counts.data_ptr())
        _, st = owners[0].histogram()
        assert st["n_dropped"] == hashed.size - 100 * n_parts
    finally:
        for o in owners:
            o.close()


@pytest.mark.parametrize("n_parts", [2, 3, 4])
def test_fused_owner_routing_on_one_device(kco, lib, n_parts):
    """the fused several-GPU form with every owner's table on this device: each context counts
    its share of the reads and adds every k-mer straight to the owner's table"""
    k = 27
    reads = reads_case(14, repeat=6)
    want, n_inst, n_dist = kco.count_reads(reads, k)
    cs = [kcgpu.Counter(k, 1 << 20, block_bytes=1 << 18) for _ in range(n_parts)]
    try:
        tables = [c.table()[0] for c in cs]
        for i, c in enumerate(cs):
            c.set_owners(i, tables)
        for j, r in enumerate(reads):
            cs[j % n_parts].add_read(r)
        for c in cs:
            c.sync()
        total = np.zeros(256, dtype=np.uint64)
        inst = dist = 0
        for c in cs:
            h, st = c.histogram()
            total += h
            inst += st["n_kmers"]
            dist += st["n_distinct"]
        assert np.array_equal(total, want) and inst == n_inst and dist == n_dist
    finally:
        for c in cs:
            c.close()


def test_linked_devices(kco, lib):
    n = kcgpu.load_library().kcgpu_device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    n = min(n, 8)
    k = 31
    reads = reads_case(15, repeat=6)
    want, n_inst, n_dist = kco.count_reads(reads, k)
    cs = [kcgpu.Counter(k, 1 << 20, block_bytes=1 << 18, device=i) for i in range(n)]
    try:
        kcgpu.link(cs)
        for j, r in enumerate(reads):
            cs[j % n].add_read(r)
        total = np.zeros(256, dtype=np.uint64)
        for c in cs:
            h, _ = c.histogram()  # waits for every linked context
            total += h
        assert np.array_equal(total, want)
    finally:
        for c in cs:
            c.close()


CHILD = r"""
import json, sys
sys.path.insert(0, {tests!r})
import numpy as np
import util
from util import kcgpu
k = {k}
reads = [bytes.fromhex(l) for l in open({reads!r}).read().split()]
c = kcgpu.Counter(k, 1 << 20, block_bytes=1 << 18)
print(c.ipc_export().hex(), flush=True)
peer = c.ipc_open(bytes.fromhex(sys.stdin.readline().strip()))
c.set_owners(1, [peer, None])
for r in reads[1::2]:
    c.add_read(r)
c.sync()
print("counted", flush=True)
assert sys.stdin.readline().strip() == "hist"
h, st = c.histogram()
print(json.dumps({{"hist": h.tolist(), "n_kmers": st["n_kmers"]}}), flush=True)
sys.stdin.readline()
c.close()
"""


def test_two_processes_share_tables_over_ipc(kco, lib, tmp_path):
    """one process per owner, as under torchrun: each maps the other's table through a CUDA IPC
    handle and its kernel adds to it directly"""
    k = 21
    reads = reads_case(16, n=3000, repeat=4)
    want, n_inst, _ = kco.count_reads(reads, k)
    rf = tmp_path / "reads.hex"
    rf.write_text("\n".join(r.hex() if r else "00" for r in reads))  # "00" = one non-base byte: dropped (shorter than k)
    reads = [bytes.fromhex(l) for l in rf.read_text().split()]
    child = subprocess.Popen([sys.executable, "-c", CHILD.format(tests=os.path.dirname(os.path.abspath(__file__)), k=k, reads=str(rf))],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    try:
        with kcgpu.Counter(k, 1 << 20, block_bytes=1 << 18) as c:
            peer = c.ipc_open(bytes.fromhex(child.stdout.readline().strip()))
            child.stdin.write(c.ipc_export().hex() + "\n")
            child.stdin.flush()
            c.set_owners(0, [None, peer])
            for r in reads[0::2]:
                c.add_read(r)
            c.sync()
            assert child.stdout.readline().strip() == "counted"
            h0, st0 = c.histogram()
            child.stdin.write("hist\n")
            child.stdin.flush()
            other = json.loads(child.stdout.readline())
            total = h0 + np.array(other["hist"], dtype=np.uint64)
            assert np.array_equal(total, want)
            assert st0["n_kmers"] + other["n_kmers"] == n_inst
            child.stdin.write("bye\n")
            child.stdin.flush()
        assert child.wait(timeout=60) == 0
    finally:
        if child.poll() is None:
            child.kill()


@pytest.mark.parametrize("name", ["k21", "k31", "exotic"])
def test_cli_prints_the_reference_histogram(lib, name):
    """kmer-cnt_b200/kc-c4 against what the reference kc-c4 printed for the same file"""
    fq = os.path.join(util.GOLDEN, f"e2e_{name}", "reads.fq.gz")
    for k in (5, 15, 21, 28, 31):
        out = subprocess.run([KC_CLI, "-k", str(k), "-t", "2", fq], check=True, capture_output=True).stdout.decode()
        assert out == open(os.path.join(GOLDEN_KC, f"{name}.k{k}.hist")).read(), (name, k)
    # a table that starts too small is grown and the file counted again, never printed short
    env = dict(os.environ, KCGPU_TABLE_SLOTS="4096")
    r = subprocess.run([KC_CLI, "-k", "31", fq], check=True, capture_output=True, env=env)
    assert r.stdout.decode() == open(os.path.join(GOLDEN_KC, f"{name}.k31.hist")).read()
    assert b"counting again" in r.stderr
    # a request larger than the device's memory is cut down to the largest table that fits
    env = dict(os.environ, KCGPU_TABLE_SLOTS=str(1 << 36))
    r = subprocess.run([KC_CLI, "-k", "31", fq], check=True, capture_output=True, env=env)
    assert r.stdout.decode() == open(os.path.join(GOLDEN_KC, f"{name}.k31.hist")).read()


def test_cli_parallel_readers(kco, lib, tmp_path):
    """-t N: a plain FASTQ cut into slices, one producer per reader thread (and per GPU)"""
    rng = np.random.default_rng(23)
    reads = util.make_genome_reads(rng, 200000, 30000, jitter=60, junk_rate=0.002, lower_rate=0.02, repeat=6)
    fq = str(tmp_path / "r.fq")
    util.write_fastq(fq, reads)
    env = dict(os.environ, VAFGPU_SLICE_BYTES="300000")
    for k in (21, 31):
        want, _, _ = kco.count_reads(reads, k)
        for t in (1, 4, 9):
            out = subprocess.run([KC_CLI, "-k", str(k), "-t", str(t), "-b", "500000", fq], check=True, capture_output=True,
                                 env=env).stdout.decode()
            assert out == kcgpu.format_histogram(want), (k, t)


def test_cli_malformed_records_end_the_file_where_the_reference_does(kco, lib, tmp_path):
    """kc-c4.c:139-156 under kt_pipeline(3, ...): a bad FASTQ record closes the block being read,
    the third empty block ends the file; where that is depends on -b"""
    rng = np.random.default_rng(8)
    reads = util.make_genome_reads(rng, 20000, 600, n_rate=0)
    fq = str(tmp_path / "bad.fq")
    for bad in ({10, 22, 34}, {57, 59, 300, 302, 304}):
        with open(fq, "wb") as fh:
            for i, r in enumerate(reads):
                fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * (len(r) - 3 if i in bad else len(r))))
        for block in (1500, 1000, 10_000_000):
            want, _, _ = kco.count_file(fq, 21, block)
            for t in (1, 4):
                out = subprocess.run([KC_CLI, "-k", "21", "-t", str(t), "-b", str(block), fq], check=True, capture_output=True,
                                     env=dict(os.environ, VAFGPU_SLICE_BYTES="20000")).stdout.decode()
                assert out == kcgpu.format_histogram(want), (sorted(bad), block, t)


def test_cli_usage_and_errors(lib, tmp_path):
    r = subprocess.run([KC_CLI], capture_output=True)
    assert r.returncode == 1 and r.stderr.startswith(b"Usage: kc-c4 [options] <in.fa>\n")
    r = subprocess.run([KC_CLI, "-p", "9", "x.fa"], capture_output=True)
    assert r.returncode == 1 and r.stderr == b"ERROR: -p should be at least 10\n"
    r = subprocess.run([KC_CLI, str(tmp_path / "missing.fa")], capture_output=True)
    assert r.returncode == 1 and r.stdout == b""


def test_properties_at_size(torch_cuda):
    """10 M reads x 150 bp generated on the device (too large for the CPU oracle): the number of
    k-mer instances equals the number of positions that end a run of k bases, the histogram
    accounts for every instance, and a second pass doubles every count"""
    torch = torch_cuda
    k, n_reads, rl = 31, 10_000_000, 150
    g = torch.Generator(device="cuda").manual_seed(1)
    genome = torch.randint(0, 4, (50_000_000,), device="cuda", dtype=torch.uint8, generator=g)
    lut = torch.tensor(list(b"ACGT"), device="cuda", dtype=torch.uint8)
    starts = torch.randint(0, genome.numel() - rl, (n_reads,), device="cuda", generator=g)
    stream = torch.empty((n_reads, rl + 1), device="cuda", dtype=torch.uint8)
    step = 2_000_000
    for lo in range(0, n_reads, step):
        idx = starts[lo:lo + step, None] + torch.arange(rl, device="cuda")[None, :]
        stream[lo:lo + step, :rl] = lut[genome[idx].long()]
    del idx
    noise = torch.rand((n_reads, rl), device="cuda", generator=g)
    stream[:, :rl][noise < 0.002] = ord("N")
    stream[:, rl] = ord("\n")
    del noise
    flat = stream.view(-1)
    assert flat.numel() % 16 == 0
    # expected instances: positions whose last k bytes are all bases, per read
    isb = (stream[:, :rl] != ord("N")).to(torch.int32)
    cs = torch.cumsum(isb, dim=1)
    win = cs[:, k - 1:] - torch.cat([torch.zeros((n_reads, 1), device="cuda", dtype=cs.dtype), cs[:, :rl - k]], dim=1)
    expect = int((win == k).sum().item())
    del isb, cs, win
    with kcgpu.Counter(k, 1 << 28) as c:
        c.count_device(flat.data_ptr(), flat.numel())
        h1, st = c.histogram()
        assert st["n_kmers"] == expect and st["n_overflow"] == 0
        assert int(h1.sum()) == st["n_distinct"]
        below = sum(int(h1[i]) * i for i in range(1, 255))
        assert below <= expect and (h1[255] > 0 or below == expect)
        c.count_device(flat.data_ptr(), flat.numel())
        h2, st2 = c.histogram()
        assert st2["n_kmers"] == 2 * expect and st2["n_distinct"] == st["n_distinct"]
        for cnt in range(1, 127):
            assert h2[2 * cnt] == h1[cnt] and h2[2 * cnt - 1] == 0
