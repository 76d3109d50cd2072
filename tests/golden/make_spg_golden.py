"""Writes tests/golden/spg/<case>.patterns.txt: what the UNMODIFIED reference snp-pattern-gen
(oracle/_ref, built from /root/reference) prints for the cases tests/util.py:make_spg_case
generates.  Only the outputs are committed; the inputs are regenerated from the seed."""
import os
import subprocess
import sys
import tempfile

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(here))
import util  # noqa: E402

CASES = [(1, 21), (2, 15), (3, 31), (4, 5), (5, 1)]

if __name__ == "__main__":
    out = os.path.join(here, "spg")
    os.makedirs(out, exist_ok=True)
    ref = os.path.join(util.REF_DIR, "snp-pattern-gen")
    for seed, k in CASES:
        fa, bed = util.make_spg_case(seed, k)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "g.fa"), "wb").write(fa)
            open(os.path.join(d, "s.bed"), "wb").write(bed)
            dst = os.path.join(out, f"seed{seed}_k{k}.patterns.txt")
            subprocess.run([ref, "-k", str(k), "-b", os.path.join(d, "s.bed"), "-f", os.path.join(d, "g.fa"), "-o", dst],
                           check=True, stderr=subprocess.DEVNULL)
            print(dst, sum(1 for _ in open(dst)), "patterns of", bed.count(b"\n"), "rows")
