"""The yak-count oracle (oracle/yak_oracle.c) against the reference: the committed outputs of the
unmodified yak-count (tests/golden/yak/, tests/golden/make_golden.sh) and, where oracle/_ref
exists, the live binary.  CPU only.  Integer work: byte-identical output."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import util
from util import kcgpu

GOLDEN_YAK = os.path.join(util.GOLDEN, "yak")
YAK_ORACLE = os.path.join(util.ORACLE_DIR, "yak_oracle")
REF = os.path.join(util.REF_DIR, "yak-count")

CASES = {  # golden suffix -> yak-count arguments
    "k31": ["-k", "31"], "k21": ["-k", "21"], "k21.b24": ["-k", "21", "-b", "24"],
    "k15.b19.H3": ["-k", "15", "-b", "19", "-H", "3"], "k27.b12": ["-k", "27", "-b", "12"],
}
TWO = {"k21.b30.two": ["-k", "21", "-b", "30"], "k21.b19.two": ["-k", "21", "-b", "19"], "k21.b0.two": ["-k", "21"]}


@pytest.fixture(scope="module")
def yko():
    return util.YakOracle()


def split_fastq(tmp_path, name):
    """the two overlapping halves tests/golden/make_golden.sh cuts a read set into"""
    with gzip.open(os.path.join(util.GOLDEN, f"e2e_{name}", "reads.fq.gz"), "rb") as fh:
        lines = fh.read().split(b"\n")[:-1]
    a, b = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
    open(a, "wb").write(b"\n".join(lines[:12000]) + b"\n")
    open(b, "wb").write(b"\n".join(lines[-12000:]) + b"\n")
    return a, b


@pytest.mark.parametrize("name", ["k21", "exotic"])
def test_oracle_prints_what_the_reference_printed(yko, tmp_path, name):
    fq = os.path.join(util.GOLDEN, f"e2e_{name}", "reads.fq.gz")
    for suffix, args in CASES.items():
        out = subprocess.run([YAK_ORACLE] + args + [fq], check=True, capture_output=True).stdout.decode()
        assert out == open(os.path.join(GOLDEN_YAK, f"{name}.{suffix}.hist")).read(), (name, suffix)
    a, b = split_fastq(tmp_path, name)
    for suffix, args in TWO.items():
        out = subprocess.run([YAK_ORACLE] + args + [a, b], check=True, capture_output=True).stdout.decode()
        assert out == open(os.path.join(GOLDEN_YAK, f"{name}.{suffix}.hist")).read(), (name, suffix)
    # with two files the small filter's false positives show: the fixture really exercises them
    assert open(os.path.join(GOLDEN_YAK, f"{name}.k21.b30.two.hist")).read() != open(os.path.join(GOLDEN_YAK, f"{name}.k21.b19.two.hist")).read()


def test_one_file_result_does_not_depend_on_the_filter(yko):
    """what the GPU form relies on: with one file every k-mer seen twice gets an entry whatever the
    filter lets through besides, and entries seen once are dropped: rows 2..1023 of the exact
    histogram, row 1 empty"""
    rng = np.random.default_rng(3)
    reads = util.make_genome_reads(rng, 20000, 3000, jitter=40, junk_rate=0.002, repeat=6)
    reads += [b"AC" * 1500, b"G" * 1043]             # counts beyond 1023, and exactly 1023
    exact = yko.count_reads(reads, 21)
    want = exact.copy()
    want[:2] = 0
    for bf_shift, n_hash in ((19, 4), (20, 2), (24, 4), (30, 6), (12, 4), (19, 1)):
        got = yko.count_reads(reads, 21, bf_shift=bf_shift, n_hash=n_hash)
        assert np.array_equal(got, want), (bf_shift, n_hash)
    assert exact[1] > 0 and exact[1023] > 0          # singletons exist, and the low-complexity reads saturate
    kco = util.KcOracle()
    h256, _, _ = kco.count_reads(reads, 21)
    assert np.array_equal(h256[:255], exact[:255]) and h256[255] == exact[255:].sum()   # kc-c4 is the same count, binned at 255


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/yak-count needs the reference checkout")
def test_oracle_against_the_live_reference(yko, tmp_path):
    rng = np.random.default_rng(17)
    reads = util.make_genome_reads(rng, 30000, 4000, jitter=60, junk_rate=0.003, lower_rate=0.02, repeat=5)
    a, b = str(tmp_path / "a.fa"), str(tmp_path / "b.fq")
    util.write_fastq(a, reads[:2500], fasta=True, line=70)
    util.write_fastq(b, reads[1500:])
    for args in (["-k", "31"], ["-k", "19", "-b", "22"], ["-k", "25", "-b", "19", "-H", "2", "-p", "10"], ["-k", "21", "-b", "25", "-p", "12"],
                 ["-k", "9", "-b", "19"], ["-k", "31", "-b", "19", "-K", "30000"], ["-k", "21", "-b", "15"]):
        for files in ([a], [a, b], [b, a]):
            ref = subprocess.run([REF, "-t", "3"] + args + files, check=True, capture_output=True).stdout
            got = subprocess.run([YAK_ORACLE] + args + files, check=True, capture_output=True).stdout
            assert got == ref, (args, files)


def test_bloom_insert_known_answers(yko):
    """yak_bf_insert (yak-count.c:86-104): positions h1, h1 + h2, ... in one 512-bit block, the step
    bumped when its low five bits are zero"""
    n_shift = 12                      # 8 blocks of 64 bytes
    bits = bytearray(1 << (n_shift - 3))
    buf = (util.C.c_char * len(bits)).from_buffer(bits)
    h = (5 << n_shift) | (37 << 3) | 6      # block 6, h1 = 37, h2 = 5
    assert yko.lib.yko_bf_insert(buf, n_shift, 4, h) == 0
    block = bits[6 * 64:7 * 64]
    assert [i for i in range(512) if block[i >> 3] >> (i & 7) & 1] == [37, 42, 47, 52]
    assert sum(bits) == sum(block)
    assert yko.lib.yko_bf_insert(buf, n_shift, 4, h) == 4
    h = (64 << n_shift) | (500 << 3) | 1    # h2 = 64: low five bits zero -> 65; wraps around 512
    assert yko.lib.yko_bf_insert(buf, n_shift, 3, h) == 0
    block = bits[64:128]
    assert [i for i in range(512) if block[i >> 3] >> (i & 7) & 1] == [53, 118, 500]
