"""The counting-mode oracle (oracle/kc_oracle.c) against the reference kc-c4: known answers of
its own functions (tests/golden/kat_kc.tsv, dumped from the reference's object code), the
histograms the reference binary printed for the committed read sets (tests/golden/kc/), and the
live binary when oracle/_ref exists.  Plus the host side of the counting ABI.  No GPU."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import util
from util import kcgpu

KC_REF = os.path.join(util.REF_DIR, "kc-c4")
GOLDEN_KC = os.path.join(util.GOLDEN, "kc")


@pytest.fixture(scope="module")
def kco():
    return util.KcOracle()


def kat_rows(kind):
    with open(os.path.join(util.GOLDEN, "kat_kc.tsv")) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if f[0] == kind:
                yield f[1:]


def test_base_table_matches_reference(kco):
    rows = list(kat_rows("nt4"))
    assert len(rows) == 256
    for b, code in rows:
        assert kco.lib.kco_nt4(int(b)) == int(code)


def test_hash64_matches_reference(kco, lib):
    kcgpu.load_library()
    n = 0
    for k, x, want in kat_rows("hash64"):
        assert kco.lib.kco_hash64(int(x, 16), int(k)) == int(want, 16)
        assert kcgpu.hash64(int(x, 16), int(k)) == int(want, 16)  # the library's host export
        n += 1
    assert n >= 400


def test_hash64_is_a_bijection(kco):
    for k in (1, 2, 5, 8):
        seen = {kco.lib.kco_hash64(x, k) for x in range(1 << 2 * k)}
        assert len(seen) == 1 << 2 * k and max(seen) < 1 << 2 * k


def test_hashed_kmers_match_reference(kco):
    n = 0
    for k, hexseq, cnt, vals in kat_rows("kmers"):
        seq = bytes.fromhex(hexseq)
        got = kco.hashed_kmers(seq, int(k))
        want = [int(v, 16) for v in vals.split(",")] if vals else []
        assert len(want) == int(cnt)
        assert got.tolist() == want, (k, seq)
        n += 1
    assert n >= 50


@pytest.mark.parametrize("name", ["k21", "k15", "k31", "exotic"])
@pytest.mark.parametrize("k", [5, 15, 21, 28, 31])
def test_histogram_matches_reference_golden(kco, name, k):
    hist, n_inst, n_dist = kco.count_file(os.path.join(util.GOLDEN, f"e2e_{name}", "reads.fq.gz"), k)
    want = open(os.path.join(GOLDEN_KC, f"{name}.k{k}.hist")).read()
    assert kcgpu.format_histogram(hist) == want
    assert int(hist.sum()) == n_dist and hist[0] == 0


@pytest.mark.skipif(not os.path.exists(KC_REF), reason="oracle/_ref is built from /root/reference (this container only)")
def test_histogram_matches_live_reference(kco, tmp_path):
    rng = np.random.default_rng(5)
    reads = util.make_genome_reads(rng, 30000, 3000, jitter=60, junk_rate=0.004, lower_rate=0.05, repeat=12)
    reads += [b"", b"ACGT", b"N" * 40, b"ACGTTGCA" * 3]
    for fasta, line in ((False, 0), (True, 60)):
        fn = str(tmp_path / ("r.fa" if fasta else "r.fq"))
        util.write_fastq(fn, reads, fasta=fasta, line=line)
        for k in (3, 10, 17, 31):
            for p in (10, 12):
                ref = subprocess.run([KC_REF, "-k", str(k), "-p", str(p), "-t", "3", "-b", "100000", fn],
                                     check=True, capture_output=True).stdout.decode()
                hist, _, _ = kco.count_file(fn, k)
                assert kcgpu.format_histogram(hist) == ref, (fasta, k, p)
                hist2, _, _ = kco.count_reads(reads, k)
                assert np.array_equal(hist, hist2)


def malformed_fastq(path, reads, bad_at):
    """a FASTQ whose records number bad_at[...] have a quality string that is too short (kseq.h:230)"""
    with open(path, "wb") as fh:
        for i, r in enumerate(reads):
            q = b"I" * (len(r) - 3 if i in bad_at and len(r) > 3 else len(r))
            fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, q))


@pytest.mark.skipif(not os.path.exists(KC_REF), reason="oracle/_ref is built from /root/reference (this container only)")
def test_malformed_record_closes_the_block_not_the_file(kco, tmp_path):
    """kc-c4.c:139-156: a record the reader rejects ends the block being read; the file goes on
    behind it unless that block is empty.  Where that is depends on -b."""
    rng = np.random.default_rng(8)
    reads = util.make_genome_reads(rng, 20000, 600, n_rate=0)  # 150 bases each: -b 1500 closes a block every 10 reads
    fn = str(tmp_path / "bad.fq")
    # a bad record swallows the header of the next one as quality, so it costs two reads.  Records
    # 10, 22 and 34 are the first of their block at -b 1500: three empty blocks, the three workers
    # of kt_pipeline(3, ...) retire one after the other and the file ends there; at any other
    # block size the three bad records cost nothing but themselves and their successors
    malformed_fastq(fn, reads, {10, 22, 34})
    got = {}
    for b in (1500, 1000, 30000, 10_000_000):
        ref = subprocess.run([KC_REF, "-k", "21", "-b", str(b), "-t", "2", fn], check=True, capture_output=True).stdout.decode()
        hist, n_inst, _ = kco.count_file(fn, 21, b)
        assert kcgpu.format_histogram(hist) == ref, b
        got[b] = n_inst
    assert got[1500] == 30 * 130
    assert got[1000] == got[30000] == got[10_000_000] == 594 * 130
    # bad records right behind each other (every second one, since each swallows its successor):
    # each but the first opens an empty block, whatever -b is
    malformed_fastq(fn, reads, {57, 59, 300, 302, 304})
    for b in (1000, 10_000_000):
        ref = subprocess.run([KC_REF, "-k", "21", "-b", str(b), fn], check=True, capture_output=True).stdout.decode()
        hist, n_inst, _ = kco.count_file(fn, 21, b)
        assert kcgpu.format_histogram(hist) == ref, b
        assert n_inst == (300 - 4) * 130  # 57..60 lost, then the file ends at 304: the third empty block


def test_saturation_and_top_bin(kco):
    """a k-mer seen more than 1023 times stays at 1023 (kc-c4.c:125) and lands in bin 255"""
    hist, n_inst, n_dist = kco.count_reads([b"A" * 2000], 11)
    assert n_inst == 1990 and n_dist == 1 and hist[255] == 1 and int(hist.sum()) == 1
    hist, _, _ = kco.count_reads([b"AC" * 150], 4)  # ACAC / GTGT canonical pair and CACA / TGTG
    assert int(hist.sum()) == 2


def test_abi_exports_match_header(lib):
    hdr = open(os.path.join(util.ROOT, "include", "kcgpu.h")).read()
    declared = re.findall(r"\b(kcgpu_[a-z0-9_]+)\s*\(", hdr)
    declared = [d for d in dict.fromkeys(declared)]
    assert set(declared) == set(kcgpu.EXPORTS), set(declared) ^ set(kcgpu.EXPORTS)
    kl = kcgpu.load_library()
    for name in kcgpu.EXPORTS:
        getattr(kl, name)


def test_no_gpu_fails_loudly(lib):
    """without a B200 the product must refuse, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(util.vafgpu.VafGpuError) as e:
        kcgpu.Counter(21, 1 << 16)
    assert e.value.code == util.vafgpu.ENOGPU


def test_owner_split_partitions_the_histogram(kco):
    """several owners: a k-mer belongs to hash mod n; per-owner histograms add up to the whole
    (what the several-GPU forms rely on)"""
    rng = np.random.default_rng(9)
    reads = util.make_genome_reads(rng, 20000, 1500, repeat=5)
    for k in (15, 31):
        hashed = np.concatenate([kco.hashed_kmers(r, k) for r in reads if len(r) >= k])
        whole, n_inst, n_dist = kco.count_reads(reads, k)
        assert n_inst == hashed.size
        for n in (2, 3, 8):
            own = kcgpu.owner_of(hashed, n)
            total = np.zeros(256, dtype=np.uint64)
            dist = 0
            for part in range(n):
                h, _, d = kco.count_hashed(hashed[own == part], k)
                total += h
                dist += d
            assert np.array_equal(total, whole) and dist == n_dist


# ---------------------------------------------------------------------------------------------
# the kernels' bookkeeping on the host (tests/cpu_sim/kc_sim.cpp over csrc/kcgpu_kernels.cuh)


@pytest.fixture(scope="module")
def ksim():
    return util.KcSim()


def test_header_hash_is_the_reference_hash(kco, ksim):
    for k, x, want in kat_rows("hash64"):
        assert ksim.lib.sim_kc_hash64(int(x, 16), int(k)) == int(want, 16)


def test_geometry_of_the_allocation(ksim):
    """table, lists, cursors and the inbox's cursor follow each other without overlap, the two
    halves of the list area (several owners) end where the cursors begin, and the tag of any
    2k-bit hash, plus one (0 is the free slot and an entry may have a count of 0), fits beside
    the 10-bit count once the region bits are taken off"""
    for k in range(1, 32):
        need = ksim.geometry(k, 4096, 64, 0)["need_bits"]
        assert 2 * k - need <= 53 and (need == 0 or 2 * k - need == 53)
        for table_bits in (12, 21, 27, 33, 36):
            for rb in sorted({need, min(max(need, table_bits - 23), 20), 12} & set(range(need, table_bits - 3))):
                for cap in (64, 96, 1 << 20):
                    g = ksim.geometry(k, 1 << table_bits, cap, rb)
                    assert g["lists"] == 8 << table_bits
                    assert g["cursors"] == g["lists"] + 8 * (cap << rb)
                    assert g["inbox_cursor"] == g["cursors"] + 256 * (1 << rb)
                    assert g["alloc"] == g["inbox_cursor"] + 256
                    assert g["inbox_cap"] == (cap << rb) // 2
                    assert g["lists2_end"] == g["cursors"]


@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 21, 27, 30, 31])
def test_window_extraction_is_count_seq_buf(kco, ksim, k):
    """the tile kernels' extraction (kc_extract16 of csrc/kcgpu_kernels.cuh: 48 bytes packed once, the two
    words of every k-mer read as windows of the packing, validity from a mask of the non-bases) gives, chunk by
    chunk, exactly the hashed canonical k-mers the oracle's byte loop gives, in order -- over reads of every
    length and alignment, lower case, N, U, junk bytes and reads shorter than k"""
    rng = np.random.default_rng(900 + k)
    reads = util.make_genome_reads(rng, 6000, 300, jitter=149, lower_rate=0.1, n_rate=0.02, junk_rate=0.01, repeat=6)
    reads += [b"", b"A", b"ACGTU" * 9, b"acgtn" * 20, b"T" * 64, bytes(range(32, 127)) * 2, b"ACGT" * 8 + b"\x00\xff" + b"TTGCA" * 10]
    stream = util.pack_stream_strict(reads, k)
    got = ksim.extract(k, stream)
    want = np.concatenate([kco.hashed_kmers(r, k) for r in reads if len(r) >= k] or [np.zeros(0, np.uint64)])
    assert got.size == want.size and np.array_equal(got, want)
    # the stream itself, byte for byte: what the reference's table makes of it
    want2 = kco.hashed_kmers(stream.tobytes().replace(b"\n", b"N"), k)
    assert np.array_equal(got, want2)


def test_tile_runs_land_where_their_cursor_says(ksim):
    """the word a region of a sorted tile gets (kc_tile_word) sends entry i of the tile to index
    region * stride + g + i - lbase of the list area when the whole run fits the list, and else flags it with its
    position in the list, for every place of the run in the tile, first and last regions, empty and full lists,
    cursors that ran on far past the capacity, halves of lists (stride > cap)"""
    rng = np.random.default_rng(12)
    cases = [(0, 1, 0, 0, 64, 64), (0, 8192, 0, 0, 8192, 8192), (63, 1, 8191, 4095, 64, 64), (64, 1, 0, 0, 64, 64),
             (60, 8, 100, 7, 64, 128), (1 << 40, 5, 8000, 4095, 1 << 30, 1 << 31), (0, 3, 8189, 0, 2, 64)]
    for _ in range(20000):
        cap = int(rng.choice([64, 96, 4096, 1 << 22, 1 << 33]))
        stride = cap * int(rng.choice([1, 2]))
        c = int(rng.integers(1, 200))
        lbase = int(rng.integers(0, 8192 - c + 1))
        g = int(rng.choice([0, max(cap - c, 0), max(cap - c + 1, 0), cap, int(rng.integers(0, 2 * cap + 1))]))
        cases.append((g, c, lbase, int(rng.integers(0, 4096)), cap, stride))
    for g, c, lbase, region, cap, stride in cases:
        assert ksim.lib.sim_kc_tile_run(g, c, lbase, region, cap, stride) == 0, (g, c, lbase, region, cap, stride)


@pytest.mark.parametrize("n_parts", [1, 2, 3, 8, 16])
def test_push_route_flush_on_the_host(kco, ksim, n_parts):
    """the several-owner form step by step on the host: inbox, region lists, table -- with lists
    that take everything and with lists so short that most k-mers go straight to the table"""
    rng = np.random.default_rng(40 + n_parts)
    reads = util.make_genome_reads(rng, 8000, 500, jitter=50, lower_rate=0.05, repeat=5)
    for k in (5, 21, 31):
        want, n_inst, _ = kco.count_reads(reads, k)
        stream = util.pack_stream_strict(reads, k)
        for cap in (1 << 15, 64):
            hist, lost, n_direct = ksim.count(k, n_parts, 18, cap, stream)
            assert lost == 0 and np.array_equal(hist, want), (k, cap)
            assert (n_direct > 0) == (cap == 64), (k, cap, n_direct)
