"""Exact-count parity of the LARGE-PANEL forms of the anchor kernel (strand-symmetric filter keys,
deferred two-level lookup) -- the forms the headline benchmark times -- and the committed golden
files through the GPU command line.  Needs a B200: -m gpu.

A panel selects these forms when its anchors no longer get 24 filter bits each
(csrc/vafgpu_tables.cpp: build_anchor_tables); the tests assert through vafgpu_stats that the
form under test really was the one launched.  Reference semantics being checked:
vaf-counter.c:349-427 (extraction, reset on a non-base) and :449-479 (lookup, every occurrence
counts, both strands), first-insert-wins map :198-252.  Integer work: bit-exact."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import util
from util import vafgpu

pytestmark = pytest.mark.gpu

N_LARGE = 22000      # patterns: enough to leave the small-panel form at every k


def large_case(tmp_path, oracle, seed, k, n_pat, n_reads):
    """a generated panel + reads that carry its k-mers at every alignment, with N runs, junk
    bytes, lower case, clipped and N-broken occurrences"""
    rng = np.random.default_rng(seed)
    pats = util.make_patterns(rng, n_pat, k, dup_every=997, bad_every=1013)
    reads = util.make_reads(rng, pats, k, n_reads, mean_len=140, jitter=60, plant=0.9, n_rate=0.01,
                            junk_rate=0.004, lower_rate=0.02)
    reads = [r.replace(b"NA", b"NNNN") for r in reads]          # N runs (config 4)
    # one k-mer of the panel at every offset modulo 16 and modulo the anchor stride, on both strands
    for off in range(48):
        p = pats[(off * 131) % len(pats)]
        km = p.ref_kmer.encode() if off % 2 else p.alt_kmer.encode()
        if b"N" in km:
            km = pats[0].ref_kmer.encode()
        if off % 3 == 0:
            km = util.revcomp(km)
        reads.append(b"G" * off + km + b"C" * (off % 5))
    pf = str(tmp_path / f"panel_k{k}.txt")
    util.write_patterns(pf, pats)
    want, _, n_coll = oracle.count_reads(pf, k, reads)
    keys, vals, n_coll2 = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    assert n_coll == n_coll2
    return pats, reads, want, keys, vals


def check_both_entry_points(k, keys, vals, n_pat, reads, want, expect_defer):
    torch = pytest.importorskip("torch")
    with vafgpu.Engine(k, keys, vals, n_pat, n_devices=1, block_bytes=1 << 18) as eng:
        for r in reads:
            eng.add_read(r)
        got, st = eng.finish()
        assert st["filter_canon"] == 1, st
        assert st["lookup_deferred"] == int(expect_defer), st
        assert np.array_equal(got, want)
        assert st["n_hits"] == int(want.astype(np.uint64).sum())
        # the resident entry point (what bench.py times), twice: counters add up
        eng.reset()
        stream = util.pack_stream(reads, k)
        d = torch.from_numpy(stream).cuda()
        counts = torch.zeros(2 * n_pat, dtype=torch.int32, device="cuda")
        cs = torch.cuda.current_stream().cuda_stream
        eng.count_device(d.data_ptr(), d.numel(), d_counts=counts.data_ptr(), stream=cs)
        eng.count_device(d.data_ptr(), d.numel(), d_counts=counts.data_ptr(), stream=cs)
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), 2 * want)
    # the literal recipe kernel on the same panel (its table geometry at this size)
    with vafgpu.Engine(k, keys, vals, n_pat, n_devices=1, flags=vafgpu.F_REFERENCE_RECIPE) as eng:
        for r in reads:
            eng.add_read(r)
        got, _ = eng.finish()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("k", [11, 13, 14, 15, 17, 18, 19, 21, 23, 26, 27, 29, 31])
def test_large_panel_exact_counts(tmp_path, oracle, lib, k):
    """every stride (1, 2, 4, 8, 16), static and run-time anchor lengths, canon with and
    without the deferred lookup"""
    pats, reads, want, keys, vals = large_case(tmp_path, oracle, 500 + k, k, N_LARGE, 12000)
    stride, _ = vafgpu.plan(k)
    assert want.sum() > 5000
    check_both_entry_points(k, keys, vals, len(pats), reads, want, expect_defer=stride >= 4)


def load_cfg2(tmp_path):
    pf = str(tmp_path / "cfg2_patterns.txt")
    with gzip.open(os.path.join(util.GOLDEN, "cfg2_patterns.txt.gz"), "rb") as src, open(pf, "wb") as dst:
        dst.write(src.read())
    return pf, vafgpu.load_patterns(pf)


def test_cfg2_golden_panel_exact_counts(tmp_path, oracle, lib):
    """the panel of the headline benchmark (the reference snp-pattern-gen's output for the
    NGSCheckMate GRCh38 BED, 50 first-wins collisions): the <8, canon, deferred, 14> form"""
    pf, pats = load_cfg2(tmp_path)
    assert len(pats) > 20000
    rng = np.random.default_rng(2024)
    reads = util.make_reads(rng, pats, 21, 30000, mean_len=150, jitter=0, plant=0.8, n_rate=0.005)
    reads += util.make_reads(rng, pats, 21, 4000, mean_len=150, jitter=70, plant=1.0, n_rate=0.03, junk_rate=0.01,
                             lower_rate=0.05)
    want, _, n_coll = oracle.count_reads(pf, 21, reads)
    keys, vals, n_coll2 = vafgpu.build_key_list(pats, 21)
    assert n_coll == n_coll2 == 50
    check_both_entry_points(21, keys, vals, len(pats), reads, want, expect_defer=True)


def test_large_panel_long_stream_spans_and_ranges(tmp_path, oracle, lib):
    """a resident stream long enough for full spans, a tail round and the resolver interrupting
    spans (many hits), checked by linearity against the oracle's counts of one copy"""
    torch = pytest.importorskip("torch")
    pf, pats = load_cfg2(tmp_path)
    rng = np.random.default_rng(77)
    reads = util.make_reads(rng, pats, 21, 20000, mean_len=150, jitter=10, plant=1.0, n_rate=0.002)
    want, _, _ = oracle.count_reads(pf, 21, reads)
    keys, vals, _ = vafgpu.build_key_list(pats, 21)
    one = util.pack_stream(reads, 21)
    reps = 400                                         # ~1.2 GB: > 4 spans per warp at 256 tiles
    d = torch.from_numpy(one).cuda().repeat(reps)
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1) as eng:
        counts = torch.zeros(2 * len(pats), dtype=torch.int32, device="cuda")
        eng.count_device(d.data_ptr(), d.numel(), d_counts=counts.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        _, st = eng.finish()
    assert st["lookup_deferred"] == 1
    assert np.array_equal(counts.cpu().numpy().view(np.uint32), (reps * want.astype(np.uint64)).astype(np.uint32))


E2E = sorted(d for d in os.listdir(util.GOLDEN) if d.startswith("e2e_"))


@pytest.mark.parametrize("name", E2E)
@pytest.mark.parametrize("threads", [1, 3])
def test_gpu_cli_reproduces_the_reference_golden_vaf(tmp_path, lib, name, threads):
    """this repository's vaf-counter on the committed inputs: the bytes the unmodified reference
    wrote (tests/golden/make_golden.sh), gzip input, SIMD byte rule (e2e_exotic)"""
    d = os.path.join(util.GOLDEN, name)
    k = open(os.path.join(d, "k")).read().strip()
    out = str(tmp_path / "o.vaf")
    exe = os.path.join(util.PKG, "vaf-counter")
    subprocess.run([exe, "-k", k, "-t", str(threads), "-b", "200000", "-p", os.path.join(d, "patterns.txt"), "-o", out,
                    os.path.join(d, "reads.fq.gz")], check=True, capture_output=True)
    assert open(out, "rb").read() == open(os.path.join(d, "expected.vaf"), "rb").read()


@pytest.mark.parametrize("name", E2E)
def test_gpu_api_reproduces_the_reference_golden_vaf(tmp_path, lib, name):
    d = os.path.join(util.GOLDEN, name)
    k = int(open(os.path.join(d, "k")).read())
    pats = vafgpu.load_patterns(os.path.join(d, "patterns.txt"))
    keys, vals, _ = vafgpu.build_key_list(pats, k)
    with gzip.open(os.path.join(d, "reads.fq.gz"), "rb") as fh:
        lines = fh.read().split(b"\n")
    reads = lines[1::4]
    for flags, simd, expected in ((0, True, "expected.vaf"), (vafgpu.F_STRICT_BYTES, False, "expected_scalar.vaf")):
        with vafgpu.Engine(k, keys, vals, len(pats), n_devices=1, flags=flags) as eng:
            for r in reads:
                eng.add_read(r)
            got, _ = eng.finish()
        assert vafgpu.format_vaf(pats, got) == open(os.path.join(d, expected)).read(), (name, expected)
