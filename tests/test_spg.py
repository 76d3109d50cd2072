"""snp-pattern-gen: the oracle restatement (oracle/spg_oracle.c) against the reference's golden
outputs and live binary (no GPU), and this repository's command line, whose genome scan runs on
the GPU, against both (-m gpu).  Output files are compared byte for byte."""
import os
import subprocess
import sys

import pytest

import util

sys.path.insert(0, util.GOLDEN)
from make_spg_golden import CASES  # noqa: E402

SPG_REF = os.path.join(util.REF_DIR, "snp-pattern-gen")
SPG_ORACLE = os.path.join(util.ORACLE_DIR, "spg_oracle")
SPG_CLI = os.path.join(util.PKG, "snp-pattern-gen")


def write_case(tmp_path, seed, k, **kw):
    fa, bed = util.make_spg_case(seed, k, **kw)
    f, b = str(tmp_path / "g.fa"), str(tmp_path / "s.bed")
    open(f, "wb").write(fa)
    open(b, "wb").write(bed)
    return f, b


def run(exe, k, bed, fa, out):
    return subprocess.run([exe, "-k", str(k), "-b", bed, "-f", fa, "-o", out], capture_output=True)


def golden(seed, k):
    return open(os.path.join(util.GOLDEN, "spg", f"seed{seed}_k{k}.patterns.txt"), "rb").read()


@pytest.mark.parametrize("seed,k", CASES)
def test_oracle_matches_reference_golden(tmp_path, oracle, seed, k):
    fa, bed = write_case(tmp_path, seed, k)
    out = str(tmp_path / "o.txt")
    assert run(SPG_ORACLE, k, bed, fa, out).returncode == 0
    assert open(out, "rb").read() == golden(seed, k)


@pytest.mark.skipif(not os.path.exists(SPG_REF), reason="oracle/_ref is built from /root/reference (this container only)")
def test_oracle_matches_live_reference(tmp_path, oracle):
    for seed, k in ((11, 21), (12, 7), (13, 27)):
        fa, bed = write_case(tmp_path, seed, k, n_snps=150)
        a, b = str(tmp_path / "ref.txt"), str(tmp_path / "ora.txt")
        assert run(SPG_REF, k, bed, fa, a).returncode == 0
        assert run(SPG_ORACLE, k, bed, fa, b).returncode == 0
        assert open(a, "rb").read() == open(b, "rb").read()
        assert len(open(a, "rb").read()) > 0 or k < 9


@pytest.mark.gpu
@pytest.mark.parametrize("seed,k", CASES)
def test_cli_matches_reference_golden(tmp_path, lib, seed, k):
    fa, bed = write_case(tmp_path, seed, k)
    out = str(tmp_path / "o.txt")
    r = run(SPG_CLI, k, bed, fa, out)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == golden(seed, k)
    assert b"Total SNPs: %d, Unique k-mer pairs: %d" % (open(bed, "rb").read().count(b"\n"), golden(seed, k).count(b"\n")) in r.stderr


@pytest.mark.gpu
def test_cli_matches_oracle_on_a_larger_genome(tmp_path, lib, oracle):
    """40 Mb over chromosomes longer than a staging block, gzip input, 3000 SNPs"""
    import gzip

    import numpy as np
    rng = np.random.default_rng(31)
    k = 21
    names = ["c%d" % i for i in range(5)]
    seqs = [util.ACGT[rng.integers(0, 4, n)] for n in (17_000_000, 12_000_000, 6_000_000, 4_999_999, 30)]
    seqs[1][2_000_000:2_400_000] = seqs[0][5_000_000:5_400_000]  # a 400 kb segmental duplication
    seqs[0][9_000_000:9_000_500] = ord("N")
    rows = []
    for i in range(3000):
        c = int(rng.integers(0, 4))
        pos = int(rng.integers(5_000_000, 5_400_000)) if (c == 0 and i % 5 == 0) else int(rng.integers(10, len(seqs[c]) - 10))
        ref = chr(seqs[c][pos])
        alt = "ACGT"[("ACGT".index(ref) + 1 + i % 3) % 4] if ref in "ACGT" else "A"
        rows.append(b"%s\t%d\t%d\trs%d\t%s\t%s\n" % (names[c].encode(), pos, pos + 1, i, ref.encode(), alt.encode()))
    fa, bed = str(tmp_path / "g.fa.gz"), str(tmp_path / "s.bed")
    with gzip.open(fa, "wb", compresslevel=1) as fh:
        for n, s in zip(names, seqs):
            fh.write(b">" + n.encode() + b"\n" + s.tobytes() + b"\n")
    open(bed, "wb").write(b"".join(rows))
    a, b = str(tmp_path / "cli.txt"), str(tmp_path / "ora.txt")
    r = run(SPG_CLI, k, bed, fa, a)
    assert r.returncode == 0, r.stderr
    assert run(SPG_ORACLE, k, bed, fa, b).returncode == 0
    got = open(a, "rb").read()
    assert got == open(b, "rb").read()
    assert 2000 < got.count(b"\n") < 2950  # the duplicated segment's SNPs are rejected, the rest kept


@pytest.mark.gpu
def test_cli_usage_and_errors(tmp_path, lib):
    r = subprocess.run([SPG_CLI, "-k", "20"], capture_output=True)
    assert r.returncode == 1 and r.stderr == b"Error: k must be odd\n"
    r = subprocess.run([SPG_CLI], capture_output=True)
    assert r.returncode == 1 and r.stderr.startswith(b"Usage: snp-pattern-gen -k 21 -b <snps.bed> -f <ref.fa> -o <patterns.txt>\n")
    r = run(SPG_CLI, 21, str(tmp_path / "no.bed"), str(tmp_path / "no.fa"), str(tmp_path / "o"))
    assert r.returncode == 1 and b"Error: failed to load FASTA file" in r.stderr
