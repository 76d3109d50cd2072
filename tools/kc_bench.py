#!/usr/bin/env python
"""kc_bench.py -- Gbases/s of the full k-mer counting mode (the kc-c4 path, BASELINE config 5).

    python tools/kc_bench.py [--reads N] [--k 31] [--steps K] [--warmup W] [--form fused|staged|both]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... tools/kc_bench.py --gpus N ...

Workload: --reads x 150 bp synthetic reads IN TOTAL (strong scaling: every rank holds reads/N of
them), drawn from both strands of one random genome of --genome bases with 1 % substitutions
and 0.5 % N, generated on the GPU as the stream the engine consumes.  One step = empty tables,
count every k-mer of the resident stream into the hash-partitioned tables, histogram.

  fused    kcgpu_count_device after kcgpu_set_owners: the counting kernel files every k-mer in
           its owner's region lists itself, over NVLink peer memory (CUDA IPC between the ranks);
           kcgpu_flush empties the lists into the tables, between barriers
  direct   the same without region lists (KCGPU_NO_LISTS): every k-mer is a compare-and-swap on
           its owner's table, wherever that is
  staged   kcgpu_extract_device -> torch.distributed.all_to_all_single (NCCL) -> kcgpu_insert_device,
           in chunks of the stream: the NCCL baseline the fused form is compared with
Timing: host clock around barrier + device synchronize on both sides (a step takes 0.1 s and
more; the flushes inside it run on the library's own stream).

Prints one JSON line (rank 0).  The CPU baseline is the UNMODIFIED reference kc-c4 (oracle/_ref)
on a bounded sample of the same reads; the same sample is also counted on the GPU through the
CLI-facing entry point (kcgpu_add_read) and the two histograms must be identical.
"""
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

READ_LEN = 150


def log(*a):
    print("[kc_bench]", *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads in total over all GPUs")
    ap.add_argument("--genome", type=int, default=1_000_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--form", default="fused", help="comma-separated: fused, direct, staged")
    ap.add_argument("--list-slots", type=int, default=0, help="per GPU; 0 = as large as the memory beside stream and table allows, up to one flush for the whole stream")
    ap.add_argument("--table-slots", type=int, default=0, help="per GPU; 0 = from the expected number of distinct k-mers")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sample", type=int, default=1_000_000, help="reads of the CPU baseline / parity sample")
    ap.add_argument("--chunk-mb", type=int, default=1024, help="stream bytes per extract/exchange/insert round (staged)")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    import bench as vb
    import kcgpu

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k = args.k
    n_reads = args.reads // world

    t0 = time.perf_counter()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)  # the same genome on every rank
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    genome = torch.empty(args.genome, dtype=torch.uint8, device=dev)
    for lo in range(0, args.genome, 1 << 28):
        n = min(1 << 28, args.genome - lo)
        genome[lo:lo + n] = acgt[torch.randint(0, 4, (n,), device=dev, generator=g)]
    donor = torch.cat([genome, genome])
    del genome
    stream, n_bytes = vb.make_stream(torch, donor, args.genome, n_reads, 77 + rank, dev)
    del donor
    torch.cuda.empty_cache()  # the lists are sized from what is free
    torch.cuda.synchronize()
    n_bases = n_reads * READ_LEN
    log(f"rank {rank}: {n_reads} reads ({stream.numel() / 1e9:.2f} GB) generated in {time.perf_counter() - t0:.1f} s")

    exp_kmers = args.reads * (READ_LEN - k + 1)
    slots = args.table_slots
    if not slots:
        # distinct k-mers: the genome's plus up to k new ones per substitution; tables at most ~0.7 full
        est = args.genome + int(args.reads * READ_LEN * 0.01 * k)
        slots = 1 << 20
        while slots < 1.4 * min(est, exp_kmers) / world:
            slots *= 2
    forms = args.form.split(",")
    state = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()


    def open_counter(form):
        """a fresh context per form: with region lists (fused), without (direct, staged)"""
        if "ctr" in state:
            barrier()  # nobody is still filing with a peer that is about to go away
            state["ctr"].close()
            barrier()
        list_slots = kcgpu.NO_LISTS
        if form == "fused":
            list_slots = args.list_slots
            if not list_slots:
                # as few flushes as the memory left beside stream and table allows: every flush
                # streams the whole table once, whatever the lists hold
                free_b, _ = torch.cuda.mem_get_info()
                room = int((free_b - slots * 8) * 0.85) // 8
                for n_int in range(1, 65):
                    list_slots = int(stream.numel() / n_int / 0.94) + (1 << 20)
                    if list_slots <= room:
                        break
        ctr = kcgpu.Counter(k, slots, device=local, list_slots=list_slots, block_bytes=256 << 20)
        state["ctr"] = ctr
        _, got_slots = ctr.table()
        _, st = ctr.histogram()
        log(f"rank {rank} [{form}]: table of {got_slots} slots ({got_slots * 8 / 1e9:.1f} GB), lists of {st['list_slots']} "
            f"({st['list_slots'] * 8 / 1e9:.1f} GB), a flush every {st['flush_bytes'] / 1e9:.2f} GB of stream")
        if world > 1 and form != "staged":
            handles = [None] * world
            dist.all_gather_object(handles, ctr.ipc_export())
            ctr.set_owners(rank, [None if r == rank else ctr.ipc_open(handles[r]) for r in range(world)])
            _, st = ctr.histogram()  # flush_bytes depends on the number of owners
        return ctr, got_slots, st["flush_bytes"]
    # a stream of our own: the legacy default stream has handle 0, which the C ABI reads as "the context's stream"
    work = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work)
    ts = work.cuda_stream
    assert ts != 0

    def total_hist(h):
        if world == 1:
            return h
        t = torch.from_numpy(h.astype(np.int64)).to(dev)
        dist.all_reduce(t)
        return t.cpu().numpy().astype(np.uint64)

    chunk = args.chunk_mb << 20
    rec = READ_LEN + 1
    chunk -= chunk % (rec * 16)
    kpc = chunk // rec * (READ_LEN - k + 1)  # k-mers per chunk at most
    cap = int(kpc / world * 1.1) + 4096
    if "staged" in forms:
        keys = torch.empty(world * cap, dtype=torch.int64, device=dev)
        recv = torch.empty(int(world * cap), dtype=torch.int64, device=dev)
        pcounts = torch.zeros(world, dtype=torch.int32, device=dev)

    def step_fused():
        ctr, flush_bytes = state["ctr"], state["flush_bytes"]
        if world == 1 or not flush_bytes:  # one process: the library flushes when it is due
            ctr.count_device(stream.data_ptr(), stream.numel(), ts)
            ctr.flush()
            return
        batch = flush_bytes - flush_bytes % (rec * 16)  # whole reads, 16-byte aligned
        for lo in range(0, stream.numel(), batch):
            ctr.count_device(stream.data_ptr() + lo, min(batch, stream.numel() - lo), ts)
            ctr.sync()
            dist.barrier()  # everybody has filed its k-mers
            ctr.flush()
            dist.barrier()  # every list is empty before anybody files again

    def step_e2e():
        """the same from page-locked HOST memory: H2D copies inside the timed region"""
        ctr, flush_bytes = state["ctr"], state["flush_bytes"]
        if world == 1 or not flush_bytes:
            ctr.submit_stream(host.data_ptr(), host.numel())
            ctr.flush()
            return
        batch = flush_bytes - flush_bytes % (rec * 16)
        for lo in range(0, host.numel(), batch):
            ctr.submit_stream(host.data_ptr() + lo, min(batch, host.numel() - lo))
            ctr.sync()
            dist.barrier()
            ctr.flush()
            dist.barrier()

    def step_staged():
        ctr = state["ctr"]
        for lo in range(0, stream.numel(), chunk):
            n = min(chunk, stream.numel() - lo)
            pcounts.zero_()
            ctr.extract_device(stream.data_ptr() + lo, n, world, keys.data_ptr(), cap, pcounts.data_ptr(), ts)
            if world == 1:
                cnt = int(pcounts[0].item())
                ctr.insert_device(keys.data_ptr(), cnt, 1, ts)
                continue
            send = pcounts.to(torch.int64)
            got = torch.empty_like(send)
            dist.all_to_all_single(got, send)
            s_list, r_list = send.tolist(), got.tolist()
            assert max(s_list) <= cap, "exchange lists too short"
            packed = torch.cat([keys[p * cap:p * cap + s_list[p]] for p in range(world)])
            out = recv[:sum(r_list)]
            dist.all_to_all_single(out, packed, output_split_sizes=r_list, input_split_sizes=s_list)
            ctr.insert_device(out.data_ptr(), out.numel(), world, ts)

    results = {}
    hists = {}
    for form in forms:
        ctr, slots, state["flush_bytes"] = open_counter(form)
        fn = step_staged if form == "staged" else step_fused
        times, hist_ms = [], []
        for it in range(args.warmup + args.steps):
            ctr.reset()
            barrier()
            t1 = time.perf_counter()
            fn()
            ctr.sync()
            barrier()
            ms = (time.perf_counter() - t1) * 1e3
            t1 = time.perf_counter()
            h, st = ctr.histogram()
            h = total_hist(h)
            t_hist = time.perf_counter() - t1
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if it >= args.warmup:
                times.append(ms)
                hist_ms.append(t_hist * 1e3)
        e2e = None
        if form == forms[0] and not args.no_e2e:
            host = torch.empty(stream.numel(), dtype=torch.uint8, pin_memory=True)
            host.copy_(stream)
            torch.cuda.synchronize()
            e2e_ms = []
            for it in range(1 + args.steps):
                ctr.reset()
                barrier()
                t1 = time.perf_counter()
                step_e2e()
                h_e, _ = ctr.histogram()  # the result comes back to the host inside the timed region
                barrier()
                ms_e = (time.perf_counter() - t1) * 1e3
                if world > 1:
                    t = torch.tensor([ms_e], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_e = float(t.item())
                if it:
                    e2e_ms.append(ms_e)
            assert np.array_equal(total_hist(h_e), h), "end-to-end histogram differs from the resident one"
            ms_e = sum(e2e_ms) / len(e2e_ms)
            e2e = {"value": args.reads * READ_LEN / ms_e / 1e6, "unit": "Gbases/s", "ms_per_step": ms_e,
                   "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": 256 * 8,
                   "sample": f"all {n_reads} reads per GPU from page-locked host memory through kcgpu_submit_stream + kcgpu_histogram"}
            del host
        hists[form] = h
        stt = torch.tensor([st["n_kmers"], st["n_distinct"], st["n_overflow"], st["n_dropped"], st["n_direct"]], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(stt)
        tot = stt.tolist()
        if form == "staged":
            tot[0] //= 2  # the extract and the insert kernel both count what they handle
        ms = sum(times) / len(times)
        results[form] = {
            "gbases_s": args.reads * READ_LEN / ms / 1e6, "gkmers_s": tot[0] / ms / 1e6, "ms_per_step": ms,
            "hist_ms": sum(hist_ms) / len(hist_ms), "n_kmers": tot[0], "n_distinct_claims": tot[1],
            "n_overflow": tot[2], "n_dropped": tot[3], "n_direct": tot[4], "n_flushes": st["n_flushes"], "distinct": int(h.sum()),
            "lib_kernel_ms": st["kernel_ms"], "e2e": e2e,
        }
        log(f"{form}: {results[form]}")
    for form in forms[1:]:
        assert np.array_equal(hists[forms[0]], hists[form]), f"{forms[0]} and {form} histograms differ"
    state["ctr"].close()
    del state["ctr"]

    out = None
    if rank == 0:
        best = max(results, key=lambda f: results[f]["gbases_s"])
        r = results[best]
        # algorithmic bytes: 1 byte per base read + one 8-byte slot read and written per k-mer
        alg = n_bases * world + 16 * r["n_kmers"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6544.3)) * world
        out = {
            "metric": "Gbases/s", "value": r["gbases_s"], "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8 in, u64 integer arithmetic and 64-bit atomics", "data": "synthetic",
            "config": {"workload": f"config 5: kc-c4 k={k} full k-mer counting, {args.reads} x {READ_LEN} bp synthetic reads in total "
                                   f"from a {args.genome}-base genome, hash-partitioned tables over {world} GPU(s)",
                       "k": k, "reads_total": args.reads, "table_slots_per_gpu": slots, "form": best,
                       "l2": "table and stream are far larger than L2",
                       "timing": "host clock around barrier + device synchronize"},
            "e2e": results[forms[0]]["e2e"], "forms": results,
            "roofline": {"bound": "hbm", "achieved": alg / r["ms_per_step"] / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": alg / r["ms_per_step"] / 1e6 / peak, "traffic": None,
                         "algorithmic_bytes": "1 B per base + 16 B (slot read + write) per k-mer instance",
                         "kernel": "kc_scan_tile_kernel + kc_flush_kernel (the whole job: 103 + 248 ms of 346 on config 5)"},
            "table_load": r["distinct"] / (slots * world),
        }

    # parity + CPU baseline on a bounded sample (rank 0)
    if rank == 0 and args.sample:
        ns = min(args.sample, n_reads)
        host = stream[:ns * rec].cpu().numpy().tobytes()
        reads = vb.stream_to_reads(host)
        tmp = tempfile.mkdtemp(dir=vb.shm_dir())
        fa = os.path.join(tmp, "sample.fa")
        with open(fa, "wb") as fh:
            fh.write(b"".join(b">r%d\n%s\n" % (i, s) for i, s in enumerate(reads)))
        ref = os.path.join(ROOT, "oracle", "_ref", "kc-c4")
        kind = "reference"
        if not os.path.exists(ref):
            ref, kind = os.path.join(ROOT, "oracle", "kc_oracle"), "port"
        cores = os.cpu_count() or 1
        best_t, best_threads, ref_out = None, None, None
        for th in sorted({1, 4, cores}) if kind == "reference" else [1]:
            t1 = time.perf_counter()
            ref_out = subprocess.run([ref, "-k", str(k), "-t", str(th), fa], check=True, capture_output=True).stdout.decode()
            dt = time.perf_counter() - t1
            log(f"{kind} kc-c4 -t {th}: {dt:.2f} s")
            if best_t is None or dt < best_t:
                best_t, best_threads = dt, th
        with kcgpu.Counter(k, 1 << 28, device=local) as c2:
            t1 = time.perf_counter()
            for s in reads:
                c2.add_read(s)
            h2, st2 = c2.histogram()
            t_api = time.perf_counter() - t1
        cli = os.path.join(ROOT, "kmer-cnt_b200", "kc-c4")
        t1 = time.perf_counter()
        cli_out = subprocess.run([cli, "-k", str(k), fa], check=True, capture_output=True,
                                 env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(local))).stdout.decode()
        t_cli = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": ns * READ_LEN / best_t / 1e9, "unit": "Gbases/s", "cores": best_threads, "kind": kind,
                               "sample": f"{ns} reads x {READ_LEN} bp of this workload as FASTA, whole process wall clock, best of -t 1/4/{cores}"}
        out["parity"] = {"histogram_vs_reference_on_sample": "identical" if kcgpu.format_histogram(h2) == ref_out else "DIFFERENT",
                         "cli_vs_reference_on_sample": "identical" if cli_out == ref_out else "DIFFERENT",
                         "forms_agree": (len(forms) > 1) or None}
        out["this_repo_on_sample"] = {"add_read_python_loop_gbases_s": ns * READ_LEN / t_api / 1e9,
                                      "cli_whole_process_gbases_s": ns * READ_LEN / t_cli / 1e9}
        for f in os.listdir(tmp):
            os.unlink(os.path.join(tmp, f))
        os.rmdir(tmp)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
