"""Parity of the yak-count mode (Bloom pre-filter, two passes, 1023-row histogram; include/kcgpu.h:
kcgpu_create_filtered / kcgpu_set_pass / kcgpu_histogram1024 and the yak-count command line) with
the oracle and with the outputs of the reference binary.  Needs a B200: -m gpu.

Reference: yak-count.c:71-104 (filter), :150-177 (insert), :445-456 (two passes + shrink),
:205-239,503 (histogram).  Integer work: every comparison is bit-exact."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import util
from util import kcgpu

pytestmark = pytest.mark.gpu

YAK_CLI = os.path.join(util.PKG, "yak-count")
GOLDEN_YAK = os.path.join(util.GOLDEN, "yak")


@pytest.fixture(scope="module")
def yko():
    return util.YakOracle()


def reads_case(seed, k, **kw):
    rng = np.random.default_rng(seed)
    reads = util.make_genome_reads(rng, kw.pop("genome", 40000), kw.pop("n", 6000), **kw)
    # counts beyond the 10-bit cap (AC..., and the all-A k-mer, canonical word 0, of the poly-T read),
    # exactly at it (the all-C k-mer) and one below it (the two k-mers of AGAG...)
    return reads + [b"", b"ACGT", b"N" * 64, b"AC" * 1500, b"G" * (1023 + k - 1), b"T" * 2000, (b"AG" * 2000)[:2044 + k - 1]]


def gpu_count(reads, k, *, bloom_bits=0, hashes=4, reads2=None, two_pass=None, **kw):
    two_pass = bloom_bits > 0 if two_pass is None else two_pass
    with kcgpu.Counter(k, kw.pop("slots", 1 << 21), block_bytes=1 << 17, bloom_bits=bloom_bits, bloom_hashes=hashes, **kw) as c:
        c.set_pass(kcgpu.PASS_CLAIM if two_pass else kcgpu.PASS_COUNT)
        for r in reads:
            c.add_read(r)
        if two_pass:
            c.set_pass(kcgpu.PASS_LOOKUP)
            for r in (reads if reads2 is None else reads2):
                c.add_read(r)
        return c.histogram1024(2 if two_pass else 0, 1023)


@pytest.mark.parametrize("k", [1, 5, 15, 21, 27, 31])
def test_single_pass_is_the_exact_histogram(yko, lib, k):
    """yak-count -b 0: every count up to the cap in its own row (kc-c4 folds them at 255)"""
    reads = reads_case(700 + k, k, jitter=100, junk_rate=0.004, lower_rate=0.03, n_rate=0.01, repeat=12)
    want = yko.count_reads(reads, k)
    got, st = gpu_count(reads, k)
    assert np.array_equal(got, want)
    assert st["n_overflow"] == 0 and st["n_distinct"] == int(want.sum())
    if k >= 15:
        assert want[1023] >= 2 and want[1022] >= 1


@pytest.mark.parametrize("bloom_bits,hashes", [(13, 4), (16, 1), (20, 4), (24, 6), (30, 2)])
def test_two_passes_with_a_filter(yko, lib, bloom_bits, hashes):
    """yak-count -b N on one file: entries for the k-mers the filter has seen before, counts from
    the second pass, entries seen once dropped -- whatever the filter's size (from hopelessly full
    to empty) the result is rows 2..1023 of the exact histogram"""
    for k in (21, 31):
        reads = reads_case(800 + k, k, jitter=60, junk_rate=0.003, repeat=10)
        want = yko.count_reads(reads, k, bf_shift=24)
        got, st = gpu_count(reads, k, bloom_bits=bloom_bits, hashes=hashes)
        assert np.array_equal(got, want), (k, bloom_bits)
        assert want[1] == 0 and want[2] > 0 and want[1023] > 0 and want[1022] > 0
        exact = yko.count_reads(reads, k)
        if bloom_bits >= 24:   # a filter with room keeps most of the k-mers seen once out of the table
            assert st["n_distinct"] < int(exact[2:].sum()) + int(exact[1]) // 20
        assert st["n_distinct"] >= int(exact[2:].sum())


def test_two_passes_without_a_filter_and_region_lists_of_every_size(yko, lib):
    """-b too small for a filter to come of it (yak-count.c:75,117): the first pass makes an entry
    for every k-mer; lists that flush every few kilobytes, once, or do not exist"""
    reads = reads_case(31, 27, jitter=50, repeat=8)
    want = yko.count_reads(reads, 27, bf_shift=12)
    for list_slots in (kcgpu.NO_LISTS, 1 << 10, 1 << 16, 0):
        got, st = gpu_count(reads, 27, bloom_bits=0, two_pass=True, list_slots=list_slots)
        assert np.array_equal(got, want), list_slots
    got, _ = gpu_count(reads, 27, bloom_bits=20, list_slots=kcgpu.NO_LISTS)
    assert np.array_equal(got, want)


def test_second_file(yko, lib):
    """counts from a second read set (yak-count a.fa b.fa): with a filter that has room the entries
    are the k-mers seen twice in the first file, as with the reference's; with a filter that is
    full they are a superset (false positives), never fewer"""
    rng = np.random.default_rng(5)
    reads = util.make_genome_reads(rng, 30000, 5000, jitter=40, repeat=6)
    a, b = reads[:3000], reads[2000:]
    want = yko.count_reads(a, 21, bf_shift=34, reads2=b)
    got, _ = gpu_count(a, 21, bloom_bits=30, reads2=b)
    assert np.array_equal(got, want)
    loose, _ = gpu_count(a, 21, bloom_bits=13, reads2=b)
    no_filter = yko.count_reads(a, 21, bf_shift=12, reads2=b)     # every k-mer of the first file gets an entry
    assert (loose >= want).all() and (loose <= no_filter).all()


def test_table_too_small_is_reported(lib):
    reads = reads_case(9, 31, n=3000)
    got, st = gpu_count(reads, 31, slots=4096, list_slots=kcgpu.NO_LISTS)
    assert st["n_overflow"] > 0


CLI_CASES = {"k31": ["-k", "31"], "k21": ["-k", "21"], "k21.b24": ["-k", "21", "-b", "24"],
             "k15.b19.H3": ["-k", "15", "-b", "19", "-H", "3"], "k27.b12": ["-k", "27", "-b", "12"]}


@pytest.mark.parametrize("name", ["k21", "exotic"])
def test_cli_prints_what_the_reference_printed(lib, tmp_path, name):
    """kmer-cnt_b200/yak-count against tests/golden/yak/ (written by the unmodified yak-count)"""
    fq = os.path.join(util.GOLDEN, f"e2e_{name}", "reads.fq.gz")
    for suffix, args in CLI_CASES.items():
        for t in ("1", "3"):
            r = subprocess.run([YAK_CLI, "-t", t] + args + [fq], check=True, capture_output=True)
            assert r.stdout.decode() == open(os.path.join(GOLDEN_YAK, f"{name}.{suffix}.hist")).read(), (name, suffix, t)
            assert b"distinct k-mers after shrinking" in r.stderr
    with gzip.open(fq, "rb") as fh:
        lines = fh.read().split(b"\n")[:-1]
    a, b = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
    open(a, "wb").write(b"\n".join(lines[:12000]) + b"\n")
    open(b, "wb").write(b"\n".join(lines[-12000:]) + b"\n")
    for suffix, args in (("k21.b30.two", ["-k", "21", "-b", "30"]), ("k21.b0.two", ["-k", "21"])):
        out = subprocess.run([YAK_CLI] + args + [a, b], check=True, capture_output=True).stdout.decode()
        assert out == open(os.path.join(GOLDEN_YAK, f"{name}.{suffix}.hist")).read(), (name, suffix)
    # a table that starts too small is grown and the files counted again
    r = subprocess.run([YAK_CLI, "-k", "21", "-b", "24", fq], check=True, capture_output=True, env=dict(os.environ, KCGPU_TABLE_SLOTS="4096"))
    assert r.stdout.decode() == open(os.path.join(GOLDEN_YAK, f"{name}.k21.b24.hist")).read()
    assert b"counting again" in r.stderr


def test_cli_usage_and_option_checks(lib):
    r = subprocess.run([YAK_CLI], capture_output=True)
    assert r.returncode == 1 and b"Usage: yak-count [options] <in.fa> [in.fa]" in r.stderr
    r = subprocess.run([YAK_CLI, "-p", "9", "x.fa"], capture_output=True)
    assert r.returncode == 1 and b"-p should be at least 10" in r.stderr


def test_linked_devices(yko, lib):
    """hash-partitioned tables (and filters) over every visible GPU, as the command line links them"""
    import torch
    nd = min(torch.cuda.device_count(), 4)
    if nd < 2:
        pytest.skip("needs two GPUs")
    reads = reads_case(77, 31, n=8000, jitter=40, repeat=8)
    want = yko.count_reads(reads, 31, bf_shift=24)
    ctrs = [kcgpu.Counter(31, 1 << 20, block_bytes=1 << 16, device=i, bloom_bits=22, bloom_hashes=4) for i in range(nd)]
    try:
        kcgpu.link(ctrs)
        ctrs[0].set_pass(kcgpu.PASS_CLAIM)
        for i, r in enumerate(reads):
            ctrs[i % nd].add_read(r)
        ctrs[0].set_pass(kcgpu.PASS_LOOKUP)
        for i, r in enumerate(reads):
            ctrs[(i + 1) % nd].add_read(r)
        total = np.zeros(1024, dtype=np.uint64)
        for c in ctrs:
            h, _ = c.histogram1024(2, 1023)
            total += h
    finally:
        for c in ctrs:
            c.close()
    assert np.array_equal(total, want)
