"""Parallel host ingest (kmer-cnt_b200/host/ingest.c): slicing a plain FASTQ over several reader
threads must hand the engine exactly the reads the one sequential reader does (which follows
kseq.h:192-232 of the reference, see test_oracle.py / test_host.py), and anything that is not
one chain of strictly formed four-line records must fall back to that reader.  CPU only: the
engine is replaced by a digest-collecting stub (tests/cpu_sim/ingest_stub.c)."""
import gzip
import os

import numpy as np
import pytest

import util


def fastq(rng, n, mean_len=150, jitter=40, qual_at=0.05, newline=True):
    out = []
    for i in range(n):
        l = max(1, int(mean_len + rng.integers(-jitter, jitter + 1)))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), l, p=[.245, .245, .245, .245, .02]))
        q = bytearray(rng.integers(33, 74, l, dtype=np.uint8).tobytes())
        if rng.random() < qual_at:
            q[0] = ord("@")          # a quality line may start with '@'
        if rng.random() < qual_at:
            q[0] = ord("+")
        out.append(b"@r%d some comment\n%s\n+\n%s\n" % (i, seq, bytes(q)))
    data = b"".join(out)
    return data if newline else data[:-1]


def write(tmp_path, name, data):
    p = os.path.join(tmp_path, name)
    with open(p, "wb") as fh:
        fh.write(data)
    return p


@pytest.mark.parametrize("slice_bytes", [4096, 10007, 65536])
@pytest.mark.parametrize("newline", [True, False])
def test_sliced_equals_sequential(tmp_path, slice_bytes, newline):
    rng = np.random.default_rng(slice_bytes)
    f = write(tmp_path, "a.fq", fastq(rng, 3000, newline=newline))
    want = util.stub_ingest([f], 21, 10_000_000, 1)
    got = util.stub_ingest([f], 21, 10_000_000, 4, slice_bytes=slice_bytes)
    assert want[3] == [0] and got[3][0] > 1          # really sliced
    assert got[:3] == want[:3]
    assert want[0][0] > 2500


def test_reads_longer_than_a_slice(tmp_path):
    rng = np.random.default_rng(5)
    f = write(tmp_path, "long.fq", fastq(rng, 40, mean_len=9000, jitter=3000))
    want = util.stub_ingest([f], 21, 10_000_000, 1)
    got = util.stub_ingest([f], 21, 10_000_000, 3, slice_bytes=4096)
    assert got[3][0] > 1 and got[:3] == want[:3]


IRREGULAR = {
    "crlf": lambda d: d.replace(b"\n", b"\r\n"),
    "blank_lines": lambda d: d.replace(b"\n@r100 ", b"\n\n@r100 "),
    "multi_line_seq": lambda d: d[:5000] + b"@ml\nACGTACGTACGTACGTACGTACGT\nACGTACGTACGTTTTTACGTAAAA\n+\n" + b"I" * 48 + b"\n" + d[5000:][d[5000:].index(b"\n@r") + 1:],
    "junk_between": lambda d: d.replace(b"\n@r200 ", b"\njunk line\n@r200 "),
    "fasta_record": lambda d: d.replace(b"\n@r300 ", b"\n>fa\nACGTACGTACGTACGTACGTACGTACGT\n@r300 "),
    "short_quality": lambda d: d[: d.index(b"\n@r400 ") - 5] + d[d.index(b"\n@r400 "):],
    "empty_sequence": lambda d: d.replace(b"\n@r500 ", b"\n@e\n\n+\n\n@r500 "),
    "trailing_blank": lambda d: d + b"\n\n",
}


@pytest.mark.parametrize("kind", sorted(IRREGULAR))
def test_irregular_files_fall_back_or_agree(tmp_path, kind):
    """Whatever the file looks like, -t 4 hands over what -t 1 does."""
    rng = np.random.default_rng(11)
    f = write(tmp_path, kind + ".fq", IRREGULAR[kind](fastq(rng, 2000)))
    want = util.stub_ingest([f], 21, 10_000_000, 1)
    got = util.stub_ingest([f], 21, 10_000_000, 4, slice_bytes=8192)
    assert got[:3] == want[:3], kind
    if kind != "trailing_blank":
        assert got[3] == [0], "must not be sliced"


def test_many_files_mixed(tmp_path):
    rng = np.random.default_rng(3)
    a = write(tmp_path, "a.fq", fastq(rng, 1500))
    b = os.path.join(tmp_path, "b.fq.gz")
    with gzip.open(b, "wb") as fh:
        fh.write(fastq(rng, 800))
    c = write(tmp_path, "c.fa", b">chr\n" + b"\n".join(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 60)) for _ in range(300)) + b"\n")
    d = os.path.join(tmp_path, "missing.fq")
    e = write(tmp_path, "e.fq", fastq(rng, 10))      # too small to slice
    files = [a, b, c, d, e]
    want = util.stub_ingest(files, 21, 100_000, 1)
    got = util.stub_ingest(files, 21, 100_000, 8, slice_bytes=16384)
    assert got[:3] == want[:3]
    assert got[3][0] > 1 and got[3][1:] == [0, 0, 0, 0]
    assert want[1][3] == 0 and want[1][2] == 1        # missing file skipped, FASTA is one record
