"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200: -m gpu.

Integer work: every comparison is bit-exact."""
import os
import subprocess

import numpy as np
import pytest

import util
from util import vafgpu

pytestmark = pytest.mark.gpu

MODES = [(0, "anchor"), (vafgpu.F_REFERENCE_RECIPE, "recipe")]


def run_engine(k, keys, vals, n, reads, flags, **kw):
    with vafgpu.Engine(k, keys, vals, n, n_devices=1, flags=flags, **kw) as eng:
        for r in reads:
            eng.add_read(r)
        return eng.finish()


def case(tmp_path, oracle, seed, k, n_pat, n_reads, **kw):
    rng = np.random.default_rng(seed)
    pats = util.make_patterns(rng, n_pat, k, dup_every=kw.pop("dup_every", 0), bad_every=kw.pop("bad_every", 0))
    reads = util.make_reads(rng, pats, k, n_reads, **kw)
    pf = str(tmp_path / "patterns.txt")
    util.write_patterns(pf, pats)
    want, n_kmers, _ = oracle.count_reads(pf, k, reads)
    keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    return pats, reads, pf, want, n_kmers, keys, vals


@pytest.mark.parametrize("flags,name", MODES)
@pytest.mark.parametrize("k", [1, 4, 11, 12, 13, 15, 16, 19, 21, 23, 26, 27, 31])
def test_counts_match_oracle_over_k(tmp_path, oracle, lib, k, flags, name):
    pats, reads, pf, want, n_kmers, keys, vals = case(
        tmp_path, oracle, 100 + k, k, 400, 6000, mean_len=120, jitter=80, n_rate=0.01,
        junk_rate=0.01, lower_rate=0.02, dup_every=37, bad_every=41)
    got, st = run_engine(k, keys, vals, len(pats), reads, flags, block_bytes=1 << 18)
    assert np.array_equal(got, want)
    assert st["n_hits"] == int(want.astype(np.uint64).sum())
    assert st["n_reads"] == sum(len(r) >= k for r in reads)
    assert st["n_bases"] == sum(len(r) for r in reads if len(r) >= k)
    if flags:
        assert st["n_kmers"] == n_kmers


@pytest.mark.parametrize("flags,name", MODES)
def test_high_n_content_resets_kmers(tmp_path, oracle, lib, flags, name):
    """config 4: 10 % N in runs; every run must restart the k-mer as vaf-counter.c:390-393 does."""
    for k in (15, 21, 31):
        rng = np.random.default_rng(k)
        pats = util.make_patterns(rng, 300, k)
        reads = util.make_reads(rng, pats, k, 5000, plant=0.9, n_rate=0.03)
        reads = [r.replace(b"NA", b"NNNN").replace(b"NC", b"NN") for r in reads]
        pf = str(tmp_path / f"p{k}.txt")
        util.write_patterns(pf, pats)
        want, _, _ = oracle.count_reads(pf, k, reads)
        keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
        got, _ = run_engine(k, keys, vals, len(pats), reads, flags)
        assert np.array_equal(got, want), k
        assert want.sum() > 0


@pytest.mark.parametrize("flags,name", MODES)
def test_edge_reads(tmp_path, oracle, lib, flags, name):
    k = 21
    rng = np.random.default_rng(5)
    pats = util.make_patterns(rng, 50, k)
    ref0, alt0 = pats[0].ref_kmer.encode(), pats[0].alt_kmer.encode()
    reads = [
        b"", b"A", ref0[:20],                      # shorter than k: dropped
        ref0,                                      # exactly k
        util.revcomp(ref0),                        # reverse strand
        ref0 + ref0 + ref0,                        # same k-mer several times in one read
        ref0[:10] + b"N" + ref0[11:],              # N in the middle
        b"N" * 40 + alt0 + b"N" * 40,
        ref0.lower(), alt0.replace(b"T", b"U"),    # lower case, U
        b"ACGT" * 50 + ref0,                       # k-mer ending at the read end
        ref0 + b"ACGT" * 50,                       # k-mer starting at the read start
        # bytes whose meaning depends on the offset within the read (SSSE3 low-nibble rule):
        # Q->A, S->C, W->G, D/E->T inside full 16-byte blocks, invalid in the tail
        ref0.replace(b"A", b"Q").replace(b"C", b"S").replace(b"G", b"W").replace(b"T", b"D") + b"ACGTACGTACG",
        b"ACGTACGTACG" + ref0.replace(b"A", b"Q").replace(b"C", b"S"),
        b"\x00\x01\x02\x03" * 8 + ref0,
    ]
    # every alignment of a k-mer relative to the 16-byte chunks and to read ends
    for off in range(0, 40):
        reads.append(b"G" * off + alt0 + b"C" * (37 - off % 7))
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    want, _, _ = oracle.count_reads(pf, k, reads)
    keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    got, _ = run_engine(k, keys, vals, len(pats), reads, flags)
    assert np.array_equal(got, want)
    assert want[0] >= 5 and want[1] >= 40


@pytest.mark.parametrize("flags,name", MODES)
def test_empty_inputs(lib, flags, name):
    keys = np.zeros(0, dtype=np.uint64)
    vals = np.zeros(0, dtype=np.uint32)
    got, st = run_engine(21, keys, vals, 0, [b"ACGT" * 30], flags)     # no patterns
    assert got.size == 0 and st["n_hits"] == 0
    keys = np.array([0x11d011da169], dtype=np.uint64)
    vals = np.array([0], dtype=np.uint32)
    got, st = run_engine(21, keys, vals, 1, [], flags)                 # no reads
    assert got.tolist() == [0, 0] and st["n_blocks"] == 0


@pytest.mark.parametrize("flags,name", MODES)
def test_read_longer_than_a_block_is_cut_with_overlap(tmp_path, oracle, lib, flags, name):
    k = 21
    rng = np.random.default_rng(11)
    pats = util.make_patterns(rng, 200, k)
    reads = util.make_reads(rng, pats, k, 3, mean_len=300_000, plant=1.0)
    big = bytearray(reads[0])
    for i in range(0, len(big) - k, 997):          # a pattern k-mer every ~1 kb, at every cut alignment
        big[i:i + k] = pats[(i // 997) % len(pats)].ref_kmer.encode()
    reads[0] = bytes(big)
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    want, _, _ = oracle.count_reads(pf, k, reads)
    keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    got, st = run_engine(k, keys, vals, len(pats), reads, flags, block_bytes=4096)
    assert np.array_equal(got, want)
    assert st["n_blocks"] > 200


def test_both_kernels_agree_and_accumulate_over_finish_calls(tmp_path, oracle, lib):
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 3, 21, 1000, 20000, plant=0.7)
    half = len(reads) // 2
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1) as eng:
        for r in reads[:half]:
            eng.add_read(r)
        first, _ = eng.finish()
        for r in reads[half:]:
            eng.add_read(r)
        total, _ = eng.finish()          # counters keep accumulating (several input files)
        eng.reset()
        zero, _ = eng.finish()
    w1, _, _ = oracle.count_reads(pf, 21, reads[:half])
    assert np.array_equal(first, w1)
    assert np.array_equal(total, want)
    assert not zero.any()


def test_submit_stream_and_count_device_entry_points(tmp_path, oracle, lib):
    torch = pytest.importorskip("torch")
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 9, 21, 500, 30000, plant=0.6, jitter=30)
    stream = util.pack_stream(reads, 21)
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1, block_bytes=1 << 16) as eng:
        eng.submit_stream(stream, n_reads=len(reads), n_bases=sum(map(len, reads)))
        got, st = eng.finish()
        assert np.array_equal(got, want)
        assert st["n_blocks"] >= stream.size // (1 << 16)
        eng.reset()
        d = torch.from_numpy(stream).cuda()
        counts = torch.zeros(2 * len(pats), dtype=torch.int32, device="cuda")
        eng.count_device(d.data_ptr(), d.numel(), d_counts=counts.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
        eng.count_device(d.data_ptr(), d.numel(), d_counts=counts.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), 2 * want)   # linearity
        own, _ = eng.finish()
        assert not own.any()


def test_uint32_counters_and_vaf_text(tmp_path, oracle, lib):
    """The CLI writes the same bytes as the oracle's writer for the same counts."""
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 21, 21, 64, 4000, plant=0.9)
    fq = str(tmp_path / "r.fq")
    util.write_fastq(fq, reads)
    out = str(tmp_path / "o.vaf")
    exe = os.path.join(util.PKG, "vaf-counter")
    subprocess.run([exe, "-k", "21", "-p", pf, "-o", out, "-b", "100000", fq], check=True, capture_output=True)
    assert open(out).read() == vafgpu.format_vaf(pats, want)


@pytest.mark.parametrize("merge", ["peer", "host"])
def test_all_visible_devices_round_robin_and_merge(tmp_path, oracle, lib, merge):
    """Blocks are dealt round-robin over every visible GPU; their kernels add into one counter
    vector over NVLink peer memory (or into one each, summed on the host); with one GPU this is
    the plain path.  The result must not depend on the number of devices (uint32 sums commute)."""
    import torch
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 31, 21, 800, 40000, plant=0.7, jitter=25)
    flags = vafgpu.F_HOST_MERGE if merge == "host" else 0
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=0, block_bytes=1 << 16, flags=flags) as eng:
        for r in reads[: len(reads) // 2]:
            eng.add_read(r)
        first, st = eng.finish()
        for r in reads[len(reads) // 2:]:
            eng.add_read(r)
        total, st = eng.finish()      # counters keep accumulating over finish() calls
    assert st["n_devices"] == torch.cuda.device_count()
    w1, _, _ = oracle.count_reads(pf, 21, reads[: len(reads) // 2])
    assert np.array_equal(first, w1)
    assert np.array_equal(total, want)
    assert st["n_blocks"] > 50


@pytest.mark.parametrize("n_threads,n_buffers", [(4, 0), (6, 2)])
def test_parallel_producers(tmp_path, oracle, lib, n_threads, n_buffers):
    """Several reader threads, each with a producer of its own (more threads than staging blocks
    in the second case): counting is additive, so the result equals the single-producer one and
    the statistics add up."""
    import threading
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 41, 21, 600, 30000, plant=0.7, jitter=25)
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1, block_bytes=1 << 15, n_buffers=n_buffers) as eng:
        errors = []

        def work(t):
            try:
                with eng.producer() as p:
                    for r in reads[t::n_threads]:
                        p.add_read(r)
            except Exception as e:  # pragma: no cover
                errors.append(e)

        ts = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors
        got, st = eng.finish()
    assert np.array_equal(got, want)
    assert st["n_reads"] == sum(1 for r in reads if len(r) >= 21)
    assert st["n_bases"] == sum(len(r) for r in reads if len(r) >= 21)


@pytest.mark.parametrize("threads", [1, 4, 9])
def test_cli_parallel_ingest_is_byte_identical(tmp_path, oracle, lib, threads):
    """vaf-counter -t N: several reader threads over slices of a plain FASTQ, a gzip file and a
    FASTA file at once; the VAF file does not depend on N (and equals the oracle's)."""
    import gzip
    pats, reads, pf, want, _, keys, vals = case(tmp_path, oracle, 77, 21, 300, 30000, plant=0.6, jitter=30, n_rate=0.01)
    a, b, c = str(tmp_path / "a.fq"), str(tmp_path / "b.fq.gz"), str(tmp_path / "c.fa")
    util.write_fastq(a, reads[:20000])
    util.write_fastq(str(tmp_path / "b.fq"), reads[20000:26000])
    with open(str(tmp_path / "b.fq"), "rb") as fi, gzip.open(b, "wb") as fo:
        fo.write(fi.read())
    util.write_fastq(c, reads[26000:], fasta=True, line=60)
    out = str(tmp_path / "o.vaf")
    exe = os.path.join(util.PKG, "vaf-counter")
    env = dict(os.environ, VAFGPU_SLICE_BYTES="200000")
    r = subprocess.run([exe, "-k", "21", "-t", str(threads), "-v", "-p", pf, "-o", out, "-b", "100000", a, b, c],
                       check=True, capture_output=True, env=env)
    assert open(out).read() == vafgpu.format_vaf(pats, want)
    n_ok = sum(1 for x in reads if len(x) >= 21)
    assert ("Sequences processed:   %d" % n_ok).encode() in r.stderr


@pytest.mark.parametrize("threads", [1, 4])
def test_cli_malformed_records_end_the_file_where_the_reference_does(tmp_path, oracle, lib, threads):
    """a FASTQ record with a short quality string closes the block being read; three of them at
    block starts end the file (the three workers of kt_pipeline, kthread.c:97-125)"""
    rng = np.random.default_rng(12)
    pats = util.make_patterns(rng, 50, 21)
    reads = util.make_reads(rng, pats, 21, 400, plant=0.9, n_rate=0)
    pf, fq = str(tmp_path / "p.txt"), str(tmp_path / "bad.fq")
    util.write_patterns(pf, pats)
    for bad in ({10, 22, 34}, {57, 59, 300, 302, 304}):
        with open(fq, "wb") as fh:
            for i, r in enumerate(reads):
                fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * (len(r) - 3 if i in bad else len(r))))
        for block in ("1500", "1000", "10000000"):
            a, b = str(tmp_path / "cli.vaf"), str(tmp_path / "orc.vaf")
            subprocess.run([os.path.join(util.PKG, "vaf-counter"), "-k", "21", "-t", str(threads), "-b", block, "-p", pf, "-o", a, fq],
                           check=True, capture_output=True, env=dict(os.environ, VAFGPU_SLICE_BYTES="20000"))
            subprocess.run([os.path.join(util.ORACLE_DIR, "vaf_oracle"), "-k", "21", "-t", "1", "-b", block, "-p", pf, "-o", b, fq],
                           check=True, capture_output=True)
            assert open(a, "rb").read() == open(b, "rb").read(), (sorted(bad), block)
