"""The oracle (oracle/vaf_oracle.c) against the reference: committed known answers and
end-to-end .vaf files produced by the unmodified reference tools (tests/golden/make_golden.sh),
and, where oracle/_ref exists (the build container), the reference binaries run live."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import util

KAT = os.path.join(util.GOLDEN, "kat_vaf.tsv")
E2E = sorted(d for d in os.listdir(util.GOLDEN) if d.startswith("e2e_"))
HAVE_REF = os.path.exists(os.path.join(util.REF_DIR, "vaf-counter"))


def kat(kind):
    with open(KAT) as fh:
        return [l.rstrip("\n").split("\t") for l in fh if l.startswith(kind + "\t")]


def test_byte_classes_match_reference_tables(oracle):
    rows = kat("nt4")
    assert len(rows) == 256
    for _, b, strict, nibble in rows:
        assert oracle.lib.vo_nt4_strict(int(b)) == int(strict), b
        assert oracle.lib.vo_nt4_nibble(int(b)) == int(nibble), b


def test_kmer_arithmetic_matches_reference(oracle):
    rows = kat("kmer")
    assert len(rows) > 200
    for r in rows:
        k, s = int(r[1]), r[2].encode()
        f = oracle.lib.vo_encode_kmer(s, k)
        if r[3] == "invalid":
            assert f == 2**64 - 1
            continue
        assert f == int(r[3], 16)
        assert oracle.lib.vo_revcomp(f, k) == int(r[4], 16)
        c = oracle.lib.vo_canonical(f, k)
        assert c == int(r[5], 16)
        assert oracle.lib.vo_kmer_hash(c) == int(r[6], 16)
        assert util.vafgpu.canonical_kmer(r[2], k) == c          # host mirror used by the CLI path


def test_bucket_function_matches_khashl(oracle):
    for _, h, bits, bucket in kat("h2b"):
        assert oracle.lib.vo_h2b(int(h, 16), int(bits)) == int(bucket)


def test_extractor_matches_reference_on_every_length_class(oracle):
    rows = kat("extract")
    assert len(rows) > 600
    for _, k, ln, hexseq, n, kmers in rows:
        seq = bytes.fromhex(hexseq)
        assert len(seq) == int(ln)
        want = [int(x, 16) for x in kmers.split(",")] if kmers else []
        assert len(want) == int(n)
        assert oracle.extract(seq, int(k), simd=True) == want, (k, hexseq)


def test_map_geometry_and_first_insert_wins(oracle, tmp_path):
    """Replays the pattern sets ref_kat.c builds?  No: those use its private RNG.  What is pinned
    here is the geometry the reference reports for n patterns (bits) and that a duplicated
    ref k-mer keeps the first pattern's value."""
    for r in kat("map"):
        n, bits = int(r[1]), int(r[2])
        want_bits = max(2, int(np.ceil(np.log2(max(3 * n, 1)))) if n else 2)
        assert bits == want_bits, (n, bits)
    rng = np.random.default_rng(1)
    pats = util.make_patterns(rng, 200, 21, dup_every=10)
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    counts, _, ncoll = oracle.count_reads(pf, 21, [p.ref_kmer.encode() for p in pats])
    assert ncoll == 19
    for i in range(10, 200, 10):          # pattern i shares its ref k-mer with i-1: i-1 wins
        assert counts[2 * i] == 0 and counts[2 * (i - 1)] == 2


@pytest.mark.parametrize("name", E2E)
def test_oracle_cli_reproduces_reference_vaf(oracle, tmp_path, name):
    d = os.path.join(util.GOLDEN, name)
    k = open(os.path.join(d, "k")).read().strip()
    fq = str(tmp_path / "reads.fq")
    with gzip.open(os.path.join(d, "reads.fq.gz")) as src, open(fq, "wb") as dst:
        shutil.copyfileobj(src, dst)
    exe = os.path.join(util.ORACLE_DIR, "vaf_oracle")
    for simd, want in (("1", "expected.vaf"), ("0", "expected_scalar.vaf")):
        out = str(tmp_path / f"o{simd}.vaf")
        subprocess.run([exe, "-k", k, "-t", "2", "-S", simd, "-p", os.path.join(d, "patterns.txt"), "-o", out,
                        os.path.join(d, "reads.fq.gz")], check=True)
        assert open(out, "rb").read() == open(os.path.join(d, want), "rb").read(), (name, simd)
    # gz and plain input, one or several files (counts accumulate over files)
    out2 = str(tmp_path / "twice.vaf")
    subprocess.run([exe, "-k", k, "-p", os.path.join(d, "patterns.txt"), "-o", out2, fq, fq], check=True)
    once = [l.split("\t") for l in open(os.path.join(d, "expected.vaf")).read().splitlines()[2:]]
    twice = [l.split("\t") for l in open(out2).read().splitlines()[2:]]
    assert all(int(b[5]) == 2 * int(a[5]) and int(b[6]) == 2 * int(a[6]) for a, b in zip(once, twice))


EDGE_FILES = {
    "multiline.fa": b">a desc\nACGTACGTAC\nGTACGTACGTACGTACG\n\nTTTT\n>b\nACGTNNACGTACGTACGTACGTACGTACGTAC\n",
    "crlf.fq": b"@r1\r\nACGTACGTACGTACGTACGTACGTA\r\n+\r\nIIIIIIIIIIIIIIIIIIIIIIIII\r\n@r2\r\nTTTTACGTACGTACGTACGTACGTACG\r\n+\r\nIIIIIIIIIIIIIIIIIIIIIIIIIII\r\n",
    "empty.fq": b"",
    "noeol.fq": b"@r1\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII",
    "truncqual.fq": b"@r1\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n@r2\nACGTACGTACGTACGTACGTACGTACGT\n+\nIII\n@r3\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "noqual.fq": b"@r1\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n@r2\nACGTACGTACGTACGTACGTACGTACGT\n+",
    "short.fq": b"@r1\nACGT\n+\nIIII\n@r2\nACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIII\n",
    "at_in_qual.fq": b"@r1\nACGTACGTACGTACGTACGTACGTA\n+\n@IIIIIIIIIIIIIIIIIIIIIIII\n@r2\nACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "lower_u.fa": b">x\nacguacguacguacguacguacguacguACGU\n",
}


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref is only built where /root/reference is mounted")
@pytest.mark.parametrize("fname", sorted(EDGE_FILES))
@pytest.mark.parametrize("block", ["10000000", "30"])
def test_live_reference_agrees_on_edge_case_files(tmp_path, fname, block):
    """Parser and block rules (kseq.h:192-232, vaf-counter.c:486-517) on awkward inputs: the
    oracle CLI and the reference binary must write the same bytes."""
    pats = [util.vafgpu.Pattern("chr1", 10, 11, "rs1", "A", "C", "ACGTACGTACGTACGTACGTA", "ACGTACGTACCTACGTACGTA"),
            util.vafgpu.Pattern("chr1", 20, 21, "rs2", "T", "G", "TTTTACGTACGTACGTACGTA", "TTTTACGTACGTACGTACGTA"[:10] + "G" + "GTACGTACGT")]
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    f = str(tmp_path / fname)
    open(f, "wb").write(EDGE_FILES[fname])
    a, b = str(tmp_path / "ref.vaf"), str(tmp_path / "orc.vaf")
    subprocess.run([os.path.join(util.REF_DIR, "vaf-counter"), "-k", "21", "-t", "1", "-b", block, "-p", pf, "-o", a, f, f],
                   check=True, capture_output=True)
    subprocess.run([os.path.join(util.ORACLE_DIR, "vaf_oracle"), "-k", "21", "-t", "1", "-b", block, "-p", pf, "-o", b, f, f],
                   check=True, capture_output=True)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref is only built where /root/reference is mounted")
def test_live_reference_agrees_on_a_fresh_synthetic_case(tmp_path):
    pre = str(tmp_path / "c")
    subprocess.run([os.path.join(util.ORACLE_DIR, "synth"), "cfg", "-o", pre, "-L", "200000", "-n", "500", "-r", "30000",
                    "-s", "77", "-x", "0.003", "-j", "30"], check=True)
    subprocess.run([os.path.join(util.REF_DIR, "snp-pattern-gen"), "-k", "21", "-b", pre + ".bed", "-f", pre + ".fa", "-o", pre + ".pat"],
                   check=True, capture_output=True)
    for t in ("1", "4"):
        subprocess.run([os.path.join(util.REF_DIR, "vaf-counter"), "-k", "21", "-t", t, "-p", pre + ".pat", "-o", pre + f".ref{t}.vaf", pre + ".fq"],
                       check=True, capture_output=True)
        subprocess.run([os.path.join(util.ORACLE_DIR, "vaf_oracle"), "-k", "21", "-t", t, "-p", pre + ".pat", "-o", pre + f".orc{t}.vaf", pre + ".fq"],
                       check=True, capture_output=True)
        assert open(pre + f".ref{t}.vaf", "rb").read() == open(pre + f".orc{t}.vaf", "rb").read()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref is only built where /root/reference is mounted")
def test_live_reference_agrees_on_malformed_records_at_block_starts(tmp_path):
    """kthread.c:97-125: a pipeline worker leaves when ITS step 0 returns NULL; the two others of
    kt_pipeline(3, ...) (vaf-counter.c:568) go on reading, so the file ends with the THIRD empty
    block.  A FASTQ record with a short quality string returns -2 (kseq.h:230), swallows the next
    header, and -- when it is the first record of a block -- makes that block empty."""
    rng = np.random.default_rng(12)
    pats = util.make_patterns(rng, 50, 21)
    reads = util.make_reads(rng, pats, 21, 400, plant=0.9, n_rate=0)  # 150 bases each: -b 1500 = 10 reads
    pf, fq = str(tmp_path / "p.txt"), str(tmp_path / "bad.fq")
    util.write_patterns(pf, pats)
    totals = {}
    for bad in ({10, 22, 34}, {10, 22}, {57, 59, 300, 302, 304}):
        with open(fq, "wb") as fh:
            for i, r in enumerate(reads):
                fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * (len(r) - 3 if i in bad else len(r))))
        for block in ("1500", "1000", "10000000"):
            a, b = str(tmp_path / "ref.vaf"), str(tmp_path / "orc.vaf")
            for exe, out in ((os.path.join(util.REF_DIR, "vaf-counter"), a), (os.path.join(util.ORACLE_DIR, "vaf_oracle"), b)):
                subprocess.run([exe, "-k", "21", "-t", "2", "-b", block, "-p", pf, "-o", out, fq], check=True, capture_output=True)
            assert open(a, "rb").read() == open(b, "rb").read(), (sorted(bad), block)
            totals[(min(bad), len(bad), block)] = sum(int(l.split("\t")[7]) for l in open(a).read().splitlines()[2:])
    # three bad block starts end the file at -b 1500 and nowhere else; two do not
    assert totals[(10, 3, "1500")] < totals[(10, 3, "1000")] == totals[(10, 3, "10000000")]
    assert totals[(10, 2, "1500")] == totals[(10, 2, "10000000")]
