#!/usr/bin/env python
"""Development: time the anchor kernel of several builds of libvafgpu.so (tools/build_variants.sh)
on one resident config-2 stream, kernel only, and check that they all count the same.

    python tools/kvar.py [--reads 20000000] [--k 21] [--patterns cfg2|<n>] lib1.so lib2.so ...

Prints one line per library: GB/s of stream (1 byte per base + separators), fraction of the
measured HBM peak.  Not a benchmark record: bench.py is."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmer-cnt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bind(path):
    lib = C.CDLL(path)
    u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    lib.vafgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, u64p, u32p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int,
                                  C.c_int, C.c_uint]
    lib.vafgpu_count_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.vafgpu_destroy.argtypes = [C.c_void_p]
    lib.vafgpu_strerror.argtypes = [C.c_void_p]
    lib.vafgpu_strerror.restype = C.c_char_p
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--patterns", default="cfg2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--n-rate", type=float, default=0.005)
    ap.add_argument("--genome", type=int, default=1 << 30)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import util
    import vafgpu
    dev = torch.device("cuda", 0)
    tmp = tempfile.mkdtemp(prefix="kvar_")
    bench.K = args.k
    if args.patterns == "cfg2":
        _, pats, keys, vals, _ = bench.load_cfg2_patterns(tmp)
    else:
        pats = util.make_patterns(np.random.default_rng(1), int(args.patterns), args.k)
        keys, vals, _ = vafgpu.build_key_list(pats, args.k)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    vals = np.ascontiguousarray(vals, dtype=np.uint32)
    donor, glen = bench.build_donor(torch, pats, args.genome, 1234, dev)
    stream, n_bytes = bench.make_stream(torch, donor, glen, args.reads, 1000, dev, n_rate=args.n_rate)
    del donor
    torch.cuda.synchronize()
    peak = 6544.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n16 = stream.numel()
    bases = args.reads * bench.READ_LEN
    first = None
    for path in args.libs:
        lib = bind(os.path.abspath(path))
        h = C.c_void_p()
        rc = lib.vafgpu_create(C.byref(h), args.k, keys.ctypes.data_as(C.POINTER(C.c_uint64)),
                               vals.ctypes.data_as(C.POINTER(C.c_uint32)), keys.size, len(pats), 1 << 20, 2, 1, 0)
        if rc:
            print(path, "create failed", rc, lib.vafgpu_strerror(None))
            continue
        counts = torch.zeros(2 * len(pats), dtype=torch.int32, device=dev)
        side = torch.cuda.Stream()          # a real stream handle: 0 would mean "the engine's own stream"
        cs = side.cuda_stream
        torch.cuda.synchronize()

        def run():
            rc = lib.vafgpu_count_device(h, 0, stream.data_ptr(), n16, counts.data_ptr(), cs)
            assert rc == 0, lib.vafgpu_strerror(h)

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        best = 1e9
        tot = 0.0
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            run()
            e1.record(side)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = min(best, ms)
            tot += ms
        counts.zero_()
        run()
        torch.cuda.synchronize()
        got = counts.cpu().numpy().copy()
        same = "first" if first is None else ("same counts" if np.array_equal(first, got) else "COUNTS DIFFER")
        if first is None:
            first = got
        avg = tot / args.steps
        print("%-40s avg %8.3f ms  %7.1f GB/s (%.3f of peak)  best %7.1f GB/s  bases/s %.1f G  hits %d  %s" % (
            os.path.basename(path), avg, bases / avg / 1e6, bases / avg / 1e6 / peak, bases / best / 1e6,
            bases / avg / 1e6, int(got.astype(np.int64).sum()), same), flush=True)
        lib.vafgpu_destroy(h)


if __name__ == "__main__":
    main()
