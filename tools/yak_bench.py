"""yak-count mode on one B200, resident stream: the two passes behind a Bloom pre-filter (first pass: entries for the
k-mers the filter has seen before; second pass: count those; yak-count.c:445-456) against the plain one-pass count of
the same reads, with the table each needs.  Parity: the 1023-row histogram of a sample against the reference binary's
(oracle/_ref/yak-count) where that exists, else against the oracle.  One JSON line on stdout.
    python tools/yak_bench.py [--reads 100000000] [--genome 1000000000] [--bloom-bits 36] [--steps 2]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmer-cnt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--genome", type=int, default=1_000_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--bloom-bits", type=int, default=36)
    ap.add_argument("--bloom-hashes", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--sample", type=int, default=300_000)
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench as vb
    import kcgpu

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (args.genome,), device=dev, generator=g)]
    stream, _ = vb.make_stream(torch, torch.cat([genome, genome]), args.genome, args.reads, 5, dev, sub_rate=0.01, n_rate=0.005)
    del genome
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    bases = args.reads * vb.READ_LEN
    est_all = args.genome + args.reads * vb.READ_LEN * 0.01 * args.k  # distinct k-mers: the genome's + the errors'
    out = {"workload": f"{args.reads} x {vb.READ_LEN} bp reads from a {args.genome}-base genome, 1 % substitutions, k = {args.k}, one B200, resident stream"}
    hists = {}
    for name, bloom in (("filtered_two_pass", args.bloom_bits), ("plain_one_pass", 0)):
        want = (args.genome * 1.6 if bloom else est_all) / 0.5
        slots = 1 << 20
        while slots < want:
            slots *= 2
        with kcgpu.Counter(args.k, slots, bloom_bits=bloom, bloom_hashes=args.bloom_hashes) as c:
            times = []
            for it in range(1 + args.steps):
                c.reset()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if bloom:
                    c.set_pass(kcgpu.PASS_CLAIM)
                    c.count_device(stream.data_ptr(), stream.numel(), s.cuda_stream)
                    c.set_pass(kcgpu.PASS_LOOKUP)
                c.count_device(stream.data_ptr(), stream.numel(), s.cuda_stream)
                h, st = c.histogram1024(2 if bloom else 0, 1023)
                ms = (time.perf_counter() - t0) * 1e3
                if it:
                    times.append(ms)
            ms = sum(times) / len(times)
            _, got_slots = c.table()
            hists[name] = h
            out[name] = {"ms": ms, "gbases_s": bases / ms / 1e6, "passes": 2 if bloom else 1, "table_gb": got_slots * 8 / 1e9,
                         "filter_gb": (1 << bloom) / 8e9 if bloom else 0.0, "entries_made": int(st["n_distinct"]),
                         "entries_kept": int(h[2:].sum()) if bloom else int(h[1:].sum()), "overflow": int(st["n_overflow"]), "flushes": int(st["n_flushes"])}
    # with one input file the filter cannot change what is kept: rows 2..1023 agree
    assert np.array_equal(hists["filtered_two_pass"][2:], hists["plain_one_pass"][2:]), "the filtered count differs from the plain one"
    out["parity"] = {"rows_2_to_1023_filtered_vs_plain": "identical"}
    # a sample against the reference binary / the oracle, through the same calls
    import util
    rec = vb.READ_LEN + 1
    reads = vb.stream_to_reads(stream[:args.sample * rec].cpu().numpy().tobytes())
    with kcgpu.Counter(args.k, 1 << 27, bloom_bits=30, bloom_hashes=4) as c:
        c.set_pass(kcgpu.PASS_CLAIM)
        for r in reads:
            c.add_read(r)
        c.set_pass(kcgpu.PASS_LOOKUP)
        for r in reads:
            c.add_read(r)
        got, _ = c.histogram1024(2, 1023)
    exe = os.path.join(ROOT, "oracle", "_ref", "yak-count")
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as d:
            fq = os.path.join(d, "s.fq")
            vb.write_fastq(fq, reads)
            txt = subprocess.run([exe, "-k", str(args.k), "-b", "30", "-t", "4", fq], check=True, capture_output=True).stdout.decode()
        ref = np.zeros(1024, dtype=np.uint64)
        for line in txt.splitlines():
            a, b = line.split()
            ref[int(a)] = int(b)
        assert np.array_equal(got[1:], ref[1:]), "sample histogram differs from the reference binary's"
        out["parity"][f"sample_{len(reads)}_reads_vs_reference_binary_b30"] = "identical"
    else:
        want = util.YakOracle().count_reads(reads, args.k, bf_shift=30)
        assert np.array_equal(got, want)
        out["parity"][f"sample_{len(reads)}_reads_vs_oracle_b30"] = "identical"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
