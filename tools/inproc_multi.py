#!/usr/bin/env python
"""The library's own several-GPU path in ONE process (what the vaf-counter command line uses):
vafgpu_create(n_devices = N), page-locked host stream dealt round-robin to the devices by
vafgpu_submit_stream, every device's kernel adding into device 0's counter vector over NVLink,
vafgpu_finish reading that one vector.  Prints end-to-end Gbases/s for N = 1 .. all visible GPUs
and checks that the counts do not depend on N.

    python tools/inproc_multi.py [--reads 40000000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmer-cnt_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=40_000_000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import vafgpu
    dev = torch.device("cuda", 0)
    tmp = tempfile.mkdtemp(prefix="inproc_")
    _, pats, keys, vals, _ = bench.load_cfg2_patterns(tmp)
    donor, glen = bench.build_donor(torch, pats, 1 << 30, 1234, dev)
    stream, n_bytes = bench.make_stream(torch, donor, glen, args.reads, 7, dev)
    del donor
    host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(stream[:n_bytes])
    del stream
    torch.cuda.synchronize()
    bases = args.reads * bench.READ_LEN
    first = None
    out = []
    for nd in range(1, torch.cuda.device_count() + 1):
        for flags, name in ((0, "peer"), (vafgpu.F_HOST_MERGE, "host-merge")):
            if nd == 1 and flags:
                continue
            with vafgpu.Engine(bench.K, keys, vals, len(pats), n_devices=nd, block_bytes=64 << 20, n_buffers=3, flags=flags) as eng:
                times = []
                for i in range(1 + args.steps):
                    eng.reset()
                    t0 = time.perf_counter()
                    eng.submit_stream((host.data_ptr(), n_bytes), n_reads=args.reads, n_bases=bases)
                    got, st = eng.finish()
                    dt = time.perf_counter() - t0
                    if i:
                        times.append(dt)
                if first is None:
                    first = got
                same = bool(np.array_equal(first, got))
                dt = sum(times) / len(times)
                out.append({"n_devices": nd, "counters": name, "gbases_s": bases / dt / 1e9, "ms": dt * 1e3, "blocks": int(st["n_blocks"]),
                            "same_counts_as_one_device": same})
                print(out[-1], flush=True)
                assert same
    print(json.dumps({"reads": args.reads, "results": out}))


if __name__ == "__main__":
    main()
