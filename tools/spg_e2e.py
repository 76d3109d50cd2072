#!/usr/bin/env python
"""End-to-end check of the snp-pattern-gen command line: the unmodified reference
(oracle/_ref/snp-pattern-gen) against this repo's (genome scan on the GPU) on a synthetic
genome of --mb megabases in 24 contigs with --snps random SNPs; byte comparison of patterns.txt,
wall clock of the whole process.  Usage: python tools/spg_e2e.py [--mb 500] [--snps 20000]"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=500)
ap.add_argument("--snps", type=int, default=20000)
ap.add_argument("--k", type=int, default=21)
a = ap.parse_args()
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
work = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
fa, bed = os.path.join(work, "g.fa"), os.path.join(work, "s.bed")
n_ctg = 24
ln = a.mb * 1_000_000 // n_ctg
rows = []
with open(fa, "wb") as fh:
    for c in range(n_ctg):
        s = acgt[rng.integers(0, 4, ln)]
        for pos in rng.integers(100, ln - 100, a.snps // n_ctg):
            ref = chr(s[pos])
            rows.append(b"chr%d\t%d\t%d\trs%d_%d\t%s\t%s\n" % (c, pos, pos + 1, c, pos, ref.encode(), "ACGT"[("ACGT".index(ref) + 1) % 4].encode()))
        fh.write(b">chr%d\n" % c)
        fh.write(s.tobytes())
        fh.write(b"\n")
open(bed, "wb").write(b"".join(rows))
print(f"== {a.mb} Mb genome in {n_ctg} contigs, {len(rows)} SNPs, k = {a.k}, host has {os.cpu_count()} cores", flush=True)
outs = {}
for name, exe in (("reference", os.path.join(ROOT, "oracle", "_ref", "snp-pattern-gen")),
                  ("this repo", os.path.join(ROOT, "kmer-cnt_b200", "snp-pattern-gen"))):
    out = os.path.join(work, name.replace(" ", "_") + ".txt")
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-k", str(a.k), "-b", bed, "-f", fa, "-o", out], capture_output=True)
    dt = time.perf_counter() - t0
    outs[name] = open(out, "rb").read() if r.returncode == 0 else None
    print(f"{name}: wall {dt * 1e3:.0f} ms, exit {r.returncode}, {0 if outs[name] is None else outs[name].count(bytes([10]))} patterns", flush=True)
print("patterns.txt", "identical" if outs["reference"] is not None and outs["reference"] == outs["this repo"] else "DIFFERENT")
for f in os.listdir(work):
    os.unlink(os.path.join(work, f))
os.rmdir(work)
