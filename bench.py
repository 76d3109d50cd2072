#!/usr/bin/env python
"""bench.py -- Gbases/s of the vaf-counter k-mer extract-and-lookup path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at N = 1: BASELINE config 2 -- k = 21, the full NGSCheckMate GRCh38 panel
(tests/golden/cfg2_patterns.txt.gz: the reference's snp-pattern-gen over a synthetic
hg38-length genome), 100 M x 150 bp synthetic reads (15 Gbases) drawn at ~5x from a diploid
donor that carries the panel's SNPs, 1 % substitutions, 0.5 % N, both strands.  The reads are
generated on the GPU (torch is plumbing: device memory, RNG, streams, torch.distributed) as the
stream the engine consumes: reads separated by '\n'.  One step = one pass of the hot path over
the whole resident stream.  N > 1: weak scaling, every rank holds its own 100 M reads, the
pattern tables are replicated, and every rank's kernel adds its hits straight into rank 0's
counter vector over NVLink (CUDA IPC; vafgpu_export_counters / vafgpu_attach_counters): the
path's only exchange is fused into the kernel and there is no collective in the step.

  value     kernel-only Gbases/s: stream already in HBM, CUDA events on the launch stream
  e2e       the same through the C ABI from page-locked HOST memory: vafgpu_submit_stream
            (H2D copies overlapped with the kernel on the engine's streams) + vafgpu_finish
            (counter read-back), wall clock around a device synchronize.  The host buffer holds
            PARSED reads (the boundary of the path): no FASTQ parsing in it
  e2e_cli   whole process against whole process on one FASTQ file in /dev/shm: this
            repository's vaf-counter and the unmodified reference's (N = 1 only)
  strong_scaling  BASELINE config 3: 600 M reads IN TOTAL sharded over the N ranks
  k_sweep   BASELINE config 4: k = 15 / 21 / 31 on reads with 7.5 % N in runs (N = 1 only)
  modes.kc  BASELINE config 5 (kc-c4 full counting) on a fixed slice (N = 1 only; tools/kc_bench.py)
  roofline  algorithmic bytes (1 byte per base) per second of the anchor kernel against the
            measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the UNMODIFIED reference vaf-counter (oracle/_ref, built from /root/reference by
            oracle/Makefile) on this box's host cores, on a bounded sample of the same reads

Nothing here runs the oracle or the reference as the product: they appear only as the
checker (parity of a sample) and as the timed CPU baseline.
"""
import argparse
import gzip
import numpy as np  # noqa: F401  (make_stream)
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "kmer-cnt_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 21
READ_LEN = 150
PATTERNS_GZ = os.path.join(ROOT, "tests", "golden", "cfg2_patterns.txt.gz")
HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
        ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chr19_KI270938v1_alt", 1066800),
        ("chr1_KI270766v1_alt", 256271)]


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload


def load_cfg2_patterns(tmpdir):
    import vafgpu
    path = os.path.join(tmpdir, "cfg2_patterns.txt")
    with gzip.open(PATTERNS_GZ, "rb") as src, open(path, "wb") as dst:
        dst.write(src.read())
    pats = vafgpu.load_patterns(path)
    keys, vals, n_coll = vafgpu.build_key_list(pats, K)
    return path, pats, keys, vals, n_coll


def build_donor(torch, pats, genome_len, seed, device, k=None):
    """Two haplotypes of a random genome of `genome_len` bases that carries every pattern's
    21-mer at its (scaled) position: 1/4 of the SNPs hom-ref, 1/2 het, 1/4 hom-alt."""
    import numpy as np
    K = k or globals()["K"]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    hap0 = torch.empty(genome_len, dtype=torch.uint8, device=device)
    step = 1 << 28
    for lo in range(0, genome_len, step):
        n = min(step, genome_len - lo)
        hap0[lo:lo + n] = acgt[torch.randint(0, 4, (n,), device=device, generator=g)]
    offs, off = {}, 0
    for name, ln in HG38:
        offs[name] = off
        off += ln
    scale = genome_len / off
    flank = K // 2
    pos = np.array([int((offs.get(p.chr, 0) + p.start) * scale) for p in pats], dtype=np.int64)
    pos = np.clip(pos, flank, genome_len - flank - 2)
    ref = np.frombuffer("".join(p.ref_kmer for p in pats).encode(), dtype=np.uint8).reshape(len(pats), K)
    alt = np.array([ord(p.alt) for p in pats], dtype=np.uint8)
    idx = torch.from_numpy(pos[:, None] - flank + np.arange(K)[None, :]).to(device)
    # later rows overwrite earlier ones where windows overlap (the 27 duplicated positions agree)
    hap0[idx.reshape(-1)] = torch.from_numpy(ref.reshape(-1).copy()).to(device)
    hap1 = hap0.clone()
    rng = np.random.default_rng(seed)
    geno = rng.integers(0, 4, len(pats))
    mid = torch.from_numpy(pos).to(device)
    alt_t = torch.from_numpy(alt).to(device)
    m1 = torch.from_numpy(geno >= 1).to(device)
    m3 = torch.from_numpy(geno == 3).to(device)
    hap1[mid[m1]] = alt_t[m1]
    hap0[mid[m3]] = alt_t[m3]
    return torch.cat([hap0, hap1]), genome_len


def make_stream(torch, donor, genome_len, n_reads, seed, device, sub_rate=0.01, n_rate=0.005, n_run_mean=0.0):
    """n_reads x 150 bp as the engine's stream: bytes + '\\n' per read, padded to 16.
    n_run_mean > 0: the N's come in runs of geometric length with that mean (config 4), n_rate
    being the fraction of bases that are N."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rec = READ_LEN + 1
    n_bytes = n_reads * rec
    stream = torch.full(((n_bytes + 15) // 16 * 16,), 10, dtype=torch.uint8, device=device)
    comp = torch.arange(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ar = torch.arange(READ_LEN, device=device)
    chunk = 1 << 20
    for lo in range(0, n_reads, chunk):
        r = min(chunk, n_reads - lo)
        start = torch.randint(0, genome_len - READ_LEN, (r,), device=device, generator=g)
        hap = torch.randint(0, 2, (r,), device=device, generator=g)
        seq = donor[(start + hap * genome_len)[:, None] + ar[None, :]]
        rev = torch.randint(0, 2, (r, 1), device=device, generator=g).bool()
        seq = torch.where(rev, comp[seq.long()].flip(1), seq)
        sub = torch.rand((r, READ_LEN), device=device, generator=g) < sub_rate * 4 / 3
        seq = torch.where(sub, acgt[torch.randint(0, 4, (r, READ_LEN), device=device, generator=g)], seq)
        if n_run_mean > 1.0:
            # a run starts with probability n_rate / mean and lasts a geometric number of bases:
            # +1 at its start, -1 after its end, running sum > 0 inside (clipped at the read end)
            start = torch.rand((r, READ_LEN), device=device, generator=g) < n_rate / n_run_mean
            u = torch.rand((r, READ_LEN), device=device, generator=g).clamp_(1e-9, 1.0)
            length = (torch.log(u) / float(np.log1p(-1.0 / n_run_mean))).floor_().to(torch.int64) + 1
            delta = torch.zeros((r, READ_LEN + 1), dtype=torch.int32, device=device)
            delta[:, :READ_LEN] += start.to(torch.int32)
            rows, cols = torch.nonzero(start, as_tuple=True)
            ends = torch.clamp(cols + length[rows, cols], max=READ_LEN)
            delta.index_put_((rows, ends), torch.full((rows.numel(),), -1, dtype=torch.int32, device=device), accumulate=True)
            isn = delta[:, :READ_LEN].cumsum(1) > 0
        else:
            isn = torch.rand((r, READ_LEN), device=device, generator=g) < n_rate
        seq = torch.where(isn, torch.full_like(seq, ord("N")), seq)
        stream[lo * rec:(lo + r) * rec].view(r, rec)[:, :READ_LEN] = seq
    return stream, n_bytes


def stream_to_reads(buf):
    """host bytes of a stream -> list of reads"""
    return [r for r in bytes(buf).split(b"\n") if r]


def write_fastq(path, reads):
    with open(path, "wb") as fh:
        q = b"I" * READ_LEN
        fh.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, q[:len(r)]) for i, r in enumerate(reads)))


def write_fastq_fixed(path, stream_bytes, n_reads):
    """n_reads fixed-length records of a host copy of the stream as a plain 4-line FASTQ file,
    assembled with numpy (a Python loop over 10^7 reads would take minutes)."""
    rec = READ_LEN + 1
    seqs = stream_bytes[:n_reads * rec].reshape(n_reads, rec)[:, :READ_LEN]
    name_w = 10                                         # "@r" + 8 digits
    row = name_w + 1 + READ_LEN + 1 + 2 + READ_LEN + 1
    chunk = 1 << 20
    with open(path, "wb") as fh:
        for lo in range(0, n_reads, chunk):
            n = min(chunk, n_reads - lo)
            out = np.empty((n, row), dtype=np.uint8)
            out[:, 0], out[:, 1] = ord("@"), ord("r")
            idx = np.arange(lo, lo + n, dtype=np.int64)
            for d in range(8):
                out[:, 2 + d] = (idx // 10 ** (7 - d)) % 10 + ord("0")
            out[:, name_w] = 10
            out[:, name_w + 1:name_w + 1 + READ_LEN] = seqs[lo:lo + n]
            at = name_w + 1 + READ_LEN
            out[:, at], out[:, at + 1], out[:, at + 2] = 10, ord("+"), 10
            out[:, at + 3:at + 3 + READ_LEN] = ord("I")
            out[:, at + 3 + READ_LEN] = 10
            fh.write(out.tobytes())


def bind_to_gpu_numa_node(torch, dev):
    """Run this process (and so allocate its page-locked buffers, first touch) on the NUMA node
    the GPU hangs off: with one process per GPU, eight ranks otherwise pin their host buffers
    wherever the launcher happened to start them.  Returns the node, None if it is not known."""
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def sweep_patterns(vafgpu, k, n, genome_len, seed):
    """n pattern rows with random k-mers at evenly spaced positions of a genome_len-base contig"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = acgt[rng.integers(0, 4, (n, k))]
    mid = k // 2
    alt_base = acgt[(np.searchsorted(acgt, ref[:, mid]) + rng.integers(1, 4, n)) % 4]
    alt = ref.copy()
    alt[:, mid] = alt_base
    step = genome_len // (n + 2)
    pats = []
    for i in range(n):
        r, a = ref[i].tobytes().decode(), alt[i].tobytes().decode()
        pats.append(vafgpu.Pattern("chrS", (i + 1) * step, (i + 1) * step + 1, "rs%d" % i, r[mid], a[mid], r, a))
    return pats


def run_k_sweep(torch, vafgpu, dev, ts, n_reads, peak):
    """BASELINE config 4: k = 15 / 21 / 31, a 1 000-SNP panel and a full-size one, reads with 7.5 %
    N in runs of geometric length (mean 5).  Kernel-only Gbases/s and fraction of the HBM roofline
    per case; the anchor kernel is checked against the literal recipe kernel on the same reads."""
    global HG38
    out = []
    glen = 1 << 28
    saved = HG38
    HG38 = [("chrS", glen)]
    try:
        for k in (15, 21, 31):
            for n_pat in (1000, 20797):
                pats = sweep_patterns(vafgpu, k, n_pat, glen, 100 + k)
                keys, vals, _ = vafgpu.build_key_list(pats, k)
                donor, _ = build_donor(torch, pats, glen, 4321, dev, k=k)
                stream, n_bytes = make_stream(torch, donor, glen, n_reads, 55 + k, dev, n_rate=0.075, n_run_mean=5.0)
                del donor
                n16 = stream.numel()
                launches = (n16 + (1 << 31) - 1) >> 31
                frac_n = float((stream[:1 << 24] == ord("N")).float().mean().item()) * (READ_LEN + 1) / READ_LEN
                with vafgpu.Engine(k, keys, vals, n_pat, n_devices=1) as eng:
                    c = torch.zeros(2 * n_pat, dtype=torch.int32, device=dev)
                    with torch.cuda.stream(ts):
                        for _ in range(3):
                            eng.count_device(stream.data_ptr(), n16, d_counts=c.data_ptr(), stream=ts.cuda_stream)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        reps = 5
                        e0.record()
                        for _ in range(reps):
                            eng.count_device(stream.data_ptr(), n16, d_counts=c.data_ptr(), stream=ts.cuda_stream)
                        e1.record()
                        torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1) / reps
                        c.zero_()
                        eng.count_device(stream.data_ptr(), n16, d_counts=c.data_ptr(), stream=ts.cuda_stream)
                        torch.cuda.synchronize()
                    _, st = eng.finish()
                with vafgpu.Engine(k, keys, vals, n_pat, n_devices=1, flags=vafgpu.F_REFERENCE_RECIPE) as rec_eng:
                    c2 = torch.zeros_like(c)
                    rec_eng.count_device(stream.data_ptr(), n16, d_counts=c2.data_ptr())
                    torch.cuda.synchronize()
                assert torch.equal(c, c2), "k sweep: anchor kernel and literal recipe kernel disagree at k=%d, %d patterns" % (k, n_pat)
                gbs = n_reads * READ_LEN / ms / 1e6
                out.append({"k": k, "patterns": n_pat, "reads": n_reads, "n_fraction": round(frac_n, 4),
                            "value": gbs, "unit": "Gbases/s", "ms_per_pass": ms, "launches_per_pass": launches,
                            "roofline_frac": gbs / peak, "hits": int(c.to(torch.int64).sum().item()),
                            "kernel": "S=%d L=%d canon=%d deferred=%d" % (st["anchor_stride"], st["anchor_len"], st["filter_canon"],
                                                                         st["lookup_deferred"]),
                            "parity": "equal to the literal recipe kernel on all %d reads" % n_reads})
                log("k sweep:", out[-1])
                del stream
                torch.cuda.empty_cache()
    finally:
        HG38 = saved
    return out


def run_kc_mode(n_reads):
    """BASELINE config 5 (kc-c4 full counting, k = 31) on a fixed slice: tools/kc_bench.py in a
    process of its own, its JSON line embedded (value, e2e, roofline, cpu_baseline, parity)."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "kc_bench.py"), "--reads", str(n_reads), "--genome", str(max(n_reads * 10, 1 << 24)),
           "--steps", "2", "--warmup", "1", "--sample", str(min(500_000, n_reads))]
    try:
        r = subprocess.run(cmd, check=True, capture_output=True, timeout=900)
        line = [l for l in r.stdout.decode().splitlines() if l.startswith("{")][-1]
        return json.loads(line)
    except Exception as exc:  # reported, not hidden
        return {"error": repr(exc)[:500]}


# ------------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (every ~2 ms: the timed region lasts
    tens of milliseconds, too short for `nvidia-smi -lms`) while the timed region runs."""

    def __init__(self, index=0):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.stop = threading.Event()
        self.th = None
        self.nv = self.h = None
        try:  # NVML start-up takes ~0.1 s: do it before the timed region, not in it
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as exc:  # no NVML: report that instead of inventing numbers
            self.reasons.add("nvml unavailable: %r" % (exc,))

    def _sample(self):
        nv, h = self.nv, self.h
        self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for k, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                       ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                       ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                       ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(k)

    def _run(self):
        try:
            while not self.stop.is_set():
                self._sample()
                time.sleep(0.001)
        except Exception as exc:
            self.reasons.add("nvml sampling failed: %r" % (exc,))

    def __enter__(self):
        if self.nv:
            self.sm.clear()
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.th:
            self.th.join(timeout=2)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline


def reference_binary():
    ref = os.path.join(ROOT, "oracle", "_ref", "vaf-counter")
    if os.path.exists(ref):
        return ref, "reference"
    port = os.path.join(ROOT, "oracle", "vaf_oracle")       # the C restatement, if the reference did not travel
    if not os.path.exists(port):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "vaf_oracle"], check=True)
    return port, "port"


def time_reference(exe, pattern_file, fastq, out, threads):
    t0 = time.perf_counter()
    subprocess.run([exe, "-k", str(K), "-t", str(threads), "-p", pattern_file, "-o", out, fastq],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def shm_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None


# ------------------------------------------------------------------------------------------------


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (config 2: 100 M)")
    ap.add_argument("--genome", type=int, default=sum(l for _, l in HG38), help="synthetic donor genome length")
    ap.add_argument("--cpu-sample-reads", type=int, default=1_000_000)
    ap.add_argument("--e2e-max-gb", type=float, default=16.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-reads", type=int, default=600_000_000, help="config 3: reads in total over all GPUs")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--e2e-cli-reads", type=int, default=10_000_000, help="reads of the whole-process comparison (0 = skip)")
    ap.add_argument("--sweep-reads", type=int, default=10_000_000, help="reads per case of the k sweep (0 = skip)")
    ap.add_argument("--kc-reads", type=int, default=20_000_000, help="reads of the counting-mode slice (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0                                    # the CPU arm runs on rank 0 alone
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys_index = local
    if world > 1 and args.impl == "ours":
        # one process per GPU: each sees exactly its own device as ordinal 0
        ids = vis.split(",") if vis else [str(i) for i in range(world)]
        os.environ["CUDA_VISIBLE_DEVICES"] = ids[local % len(ids)]
        phys_index = int(ids[local % len(ids)]) if ids[local % len(ids)].isdigit() else local
    elif vis and vis.split(",")[0].isdigit():
        phys_index = int(vis.split(",")[0])

    import numpy as np
    import torch
    import sharding
    import vafgpu

    tmp = tempfile.mkdtemp(prefix="vafbench_", dir=shm_dir())
    pattern_file, pats, keys, vals, n_coll = load_cfg2_patterns(tmp)
    n_pat = len(pats)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peaks else "fallback 6.65 TB/s"
    config = {"workload": "config 2: vaf-counter k=21, NGSCheckMate GRCh38 panel (%d patterns from a synthetic "
                          "hg38-length reference), %d x %d bp synthetic reads per GPU" % (n_pat, args.reads, READ_LEN),
              "k": K, "patterns": n_pat, "reads_per_gpu": args.reads, "read_len": READ_LEN,
              "l2": "inputs (%.1f GB per GPU) are far larger than L2" % (args.reads * (READ_LEN + 1) / 1e9),
              "parallelism": "reads sharded per rank, tables replicated, every rank's kernel adds its hits into rank 0's "
                             "counter vector over NVLink (no collective in the step; all-reduce only where peer access is missing)"}

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); --impl reference also builds its "
                         "sample of the workload on the GPU")
    dev = torch.device("cuda", 0 if world > 1 else local)
    torch.cuda.set_device(dev)

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        exe, kind = reference_binary()
        donor, glen = build_donor(torch, pats, min(args.genome, 1 << 28), 1234, dev)
        s, nb = make_stream(torch, donor, glen, args.cpu_sample_reads, 99, dev)
        reads = stream_to_reads(s[:nb].cpu().numpy().tobytes())
        fq = os.path.join(tmp, "sample.fq")
        write_fastq(fq, reads)
        bases = sum(map(len, reads))
        ncpu = os.cpu_count() or 1
        best_t = 1
        t1 = time_reference(exe, pattern_file, fq, os.path.join(tmp, "r.vaf"), 1)
        tn = time_reference(exe, pattern_file, fq, os.path.join(tmp, "r.vaf"), ncpu)
        threads = 1 if t1 <= tn else ncpu          # more threads make the reference slower
        log(f"reference: -t 1 {bases / t1 / 1e6:.1f} Mbases/s, -t {ncpu} {bases / tn / 1e6:.1f} Mbases/s")
        times = []
        for i in range(args.warmup + args.steps):
            dt = time_reference(exe, pattern_file, fq, os.path.join(tmp, "r.vaf"), threads)
            if i >= args.warmup:
                times.append(dt)
        total = sum(times)
        v = bases * len(times) / total / 1e9
        line = {"impl": "reference", "metric": "Gbases/s", "value": v, "unit": "Gbases/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": threads, "kind": kind,
                                 "sample": "%d reads x %d bp of the config-2 workload per step; whole process wall "
                                           "clock (pattern load + FASTQ parse + count); best of -t 1 (%.4f) and -t %d "
                                           "(%.4f Gbases/s)" % (len(reads), READ_LEN, bases / t1 / 1e9, ncpu, bases / tn / 1e9)},
                "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None
    numa = bind_to_gpu_numa_node(torch, dev)         # page-locked buffers then come from the GPU's own node
    t_gen = time.perf_counter()
    donor, glen = build_donor(torch, pats, args.genome, 1234, dev)
    stream, n_bytes = make_stream(torch, donor, glen, args.reads, 1000 + rank, dev)
    del donor
    torch.cuda.synchronize()
    log(f"generated {args.reads} reads ({n_bytes / 1e9:.2f} GB) in {time.perf_counter() - t_gen:.1f} s")
    bases_per_step = args.reads * READ_LEN
    rec = READ_LEN + 1

    eng = vafgpu.Engine(K, keys, vals, n_pat, n_devices=1, block_bytes=256 << 20, n_buffers=3)
    ts = torch.cuda.Stream()
    n16 = stream.numel()
    launches_per_step = (n16 + (1 << 31) - 1) >> 31

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def scan(n_bytes16, d_counts=0):
        eng.count_device(stream.data_ptr(), n_bytes16, d_counts=d_counts, stream=ts.cuda_stream)

    def new_vec():
        return torch.zeros(2 * n_pat, dtype=torch.int32, device=dev)

    # what one pass over this rank's reads counts (the expectation for everything below)
    one, scratch = new_vec(), new_vec()
    with torch.cuda.stream(ts):
        scan(n16, one.data_ptr())
    torch.cuda.synchronize()
    hits_per_step = int(one.to(torch.int64).sum().item())
    # one counter vector for all ranks: rank 0's, the others' kernels add into it over NVLink
    merge = sharding.share_counters(eng, dist, rank)
    log(f"counters: {merge}")
    shared = merge != "all_reduce"      # hits land in the engine's (shared) vector: nothing to merge
    own = None if shared else new_vec()

    def step():
        if shared:
            scan(n16)
        else:
            own.zero_()
            scan(n16, own.data_ptr())
            sharding.all_reduce_counts(own, dist)

    # strong scaling, BASELINE config 3: 600 M reads IN TOTAL over the ranks; a rank scans its
    # share as whole passes over its resident stream plus a partial one
    share = sharding.strong_share(args.strong_reads, rank, world)
    full_passes, part = divmod(share, args.reads)
    part16 = (part * rec + 15) // 16 * 16

    def strong_step():
        tgt = 0 if shared else scratch.data_ptr()
        if not shared:
            scratch.zero_()
        for _ in range(full_passes):
            scan(n16, tgt)
        if part:
            scan(part16, tgt)
        if not shared:
            sharding.all_reduce_counts(scratch, dist)

    eng.reset()
    barrier()
    with torch.cuda.stream(ts):
        for _ in range(args.warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(phys_index)
        with clk:  # clocks are sampled over the timed loops (each lasts only tens of milliseconds)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            # the dominant kernel's average launch duration, for the roofline: with the counters shared (or one
            # GPU) the timed region above holds nothing but its launches; where a step ends with an all-reduce
            # the same scans are timed once more without it
            if shared:
                kernel_ms = ms / (args.steps * launches_per_step)
                kernel_ms_source = "CUDA events around the timed region of `value` (nothing but this kernel's launches in it) / launches"
            else:
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                for _ in range(args.steps):
                    scan(n16, scratch.data_ptr())
                k1.record()
                barrier()
                kernel_ms = k0.elapsed_time(k1) / (args.steps * launches_per_step)
                kernel_ms_source = "CUDA events around the same scans without the step's all-reduce / launches" 
            scratch.zero_()
            strong_step()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.strong_steps):
                strong_step()
            s1.record()
            barrier()
            strong_ms = s0.elapsed_time(s1) / args.strong_steps
    if world > 1:
        t = torch.tensor([ms, strong_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, strong_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = ms / args.steps
    value = bases_per_step * world / (ms_per_step * 1e-3) / 1e9
    strong = {"workload": "config 3: %d x %d bp reads in total, sharded over %d GPU(s)" % (args.strong_reads, READ_LEN, world),
              "value": args.strong_reads * READ_LEN / (strong_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": strong_ms,
              "steps": args.strong_steps, "scaling": "strong", "merge": merge,
              "note": "a rank's share is scanned as passes over its resident %d-read stream" % args.reads}

    # ---- parity at full size: the merged counters, linearity over a split, the literal recipe
    #      kernel on a slice, the oracle and the reference on a sample
    parity = {}
    part_counts = new_vec()
    with torch.cuda.stream(ts):
        if part:
            scan(part16, part_counts.data_ptr())
        cut = (args.reads // 2) * rec
        cut16 = (cut + 15) // 16 * 16
        head = torch.full((cut16,), 10, dtype=torch.uint8, device=dev)
        head[:cut] = stream[:cut]
        tail = torch.full(((n_bytes - cut + 15) // 16 * 16,), 10, dtype=torch.uint8, device=dev)
        tail[:n_bytes - cut] = stream[cut:n_bytes]
        halves = new_vec()
        eng.count_device(head.data_ptr(), head.numel(), d_counts=halves.data_ptr(), stream=ts.cuda_stream)
        eng.count_device(tail.data_ptr(), tail.numel(), d_counts=halves.data_ptr(), stream=ts.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(halves, one), "linearity over a split of the stream failed"
    parity["split_linearity"] = "ok"
    del head, tail, halves
    # everything the ranks counted into the merged vector: weak passes and strong shares
    share_counts = one.to(torch.int64) * full_passes + part_counts.to(torch.int64)
    if shared:
        want = one.to(torch.int64) * (args.warmup + args.steps) + share_counts * (1 + args.strong_steps)
        barrier()                          # every rank has drained before rank 0 reads the shared vector
        got_shared, _ = eng.finish()
    else:
        want = share_counts                  # the last strong step, summed over the ranks
        got_shared = scratch.cpu().numpy().view(np.uint32)
    if world > 1:
        dist.all_reduce(want)
    if rank == 0:
        assert np.array_equal(want.cpu().numpy().astype(np.uint32), got_shared), \
            "the merged counters are not the sum of what the ranks counted"
        parity["merged_counters"] = "identical to the sum of the ranks' counts (%s)" % merge
    if world > 1:
        dist.barrier()
    _, st_kernel = eng.finish()
    slice_reads = min(args.reads, 2_000_000)
    sl = stream[:(slice_reads * rec + 15) // 16 * 16].clone()
    sl[slice_reads * rec:] = 10
    with vafgpu.Engine(K, keys, vals, n_pat, n_devices=1, flags=vafgpu.F_REFERENCE_RECIPE) as rec_eng:
        c2 = new_vec()
        rec_eng.count_device(sl.data_ptr(), sl.numel(), d_counts=c2.data_ptr())
        torch.cuda.synchronize()
    c1 = new_vec()
    eng.count_device(sl.data_ptr(), sl.numel(), d_counts=c1.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(c1, c2), "anchor kernel and literal recipe kernel disagree"
    parity["recipe_kernel_on_%d_reads" % slice_reads] = "ok"
    del sl

    # ---- end to end through the C ABI from page-locked host memory (its own engine: the shared
    #      vector stays out of it; counters are merged the way a caller without NVLink would)
    avail_gb = 0.0
    try:
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable"):
                avail_gb = int(l.split()[1]) / 1e6
    except OSError:
        pass
    e2e_bytes = int(min(n_bytes, args.e2e_max_gb * 1e9, max(avail_gb, 1.0) * 1e9 / (3 * max(1, min(world, 8)))))
    e2e_reads = e2e_bytes // rec
    e2e_bytes = e2e_reads * rec
    host = torch.empty(e2e_bytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(stream[:e2e_bytes])
    torch.cuda.synchronize()
    e2e_eng = vafgpu.Engine(K, keys, vals, n_pat, n_devices=1, block_bytes=256 << 20, n_buffers=3)
    e2e_times, got = [], None
    for i in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        e2e_eng.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_eng.submit_stream((host.data_ptr(), e2e_bytes), n_reads=e2e_reads, n_bases=e2e_reads * READ_LEN)
        got, st_e2e = e2e_eng.finish()
        if world > 1:
            g = torch.from_numpy(got.astype(np.int32)).to(dev)
            dist.all_reduce(g)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_times.append(dt)
    e2e_dt = sum(e2e_times) / len(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = e2e_reads * READ_LEN * world / e2e_dt / 1e9
    # what came back must be what the resident pass counted for the same reads
    chk = new_vec()
    n_e16 = (e2e_bytes + 15) // 16 * 16
    if n_e16 <= n16:
        part_s = stream[:n_e16].clone()
        part_s[e2e_bytes:] = 10
        eng.count_device(part_s.data_ptr(), part_s.numel(), d_counts=chk.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(chk.cpu().numpy().view(np.uint32), got), "e2e counters differ from the resident pass"
        del part_s
        parity["e2e_equals_resident"] = "ok"
    e2e_eng.close()
    del host

    # ---- CPU baseline + byte parity of the .vaf on a sample (rank 0 only)
    cpu = None
    e2e_cli = None
    cli_job = None
    if rank == 0 and not args.no_cpu_baseline:
        exe, kind = reference_binary()
        n_s = min(args.cpu_sample_reads, args.reads)
        reads = stream_to_reads(stream[:n_s * rec].cpu().numpy().tobytes())
        fq = os.path.join(tmp, "sample.fq")
        write_fastq(fq, reads)
        bases = sum(map(len, reads))
        ncpu = os.cpu_count() or 1
        t1 = time_reference(exe, pattern_file, fq, os.path.join(tmp, "ref1.vaf"), 1)
        tn = time_reference(exe, pattern_file, fq, os.path.join(tmp, "refn.vaf"), ncpu)
        best, threads = (t1, 1) if t1 <= tn else (tn, ncpu)
        cpu = {"value": bases / best / 1e9, "unit": "Gbases/s", "cores": threads, "kind": kind,
               "sample": "%d reads x %d bp of this workload, whole process wall clock; -t 1: %.4f, -t %d: %.4f Gbases/s; "
                         "host has %d cores" % (len(reads), READ_LEN, bases / t1 / 1e9, ncpu, bases / tn / 1e9, ncpu)}
        # our CLI-equivalent path on the same FASTQ sample: byte-identical .vaf
        with vafgpu.Engine(K, keys, vals, n_pat, n_devices=1) as e3:
            for r in reads:
                e3.add_read(r)
            mine, _ = e3.finish()
        ours_txt = vafgpu.format_vaf(pats, mine)
        ref_txt = open(os.path.join(tmp, "ref1.vaf")).read()
        assert ours_txt == ref_txt, "VAF text differs from the reference's on the sample"
        assert open(os.path.join(tmp, "refn.vaf")).read() == ref_txt
        parity["vaf_bytes_vs_%s_on_%d_reads" % (kind, len(reads))] = "identical"
        try:
            import util
            want_o, _, _ = util.Oracle().count_reads(pattern_file, K, reads)
            assert np.array_equal(want_o, mine)
            parity["oracle_counts_on_sample"] = "identical"
        except Exception as exc:  # the oracle is a checker; its absence is reported, not hidden
            parity["oracle_counts_on_sample"] = "not run: %r" % (exc,)
        del reads
        # ---- whole process against whole process on one FASTQ file (N = 1): the file is written now, while the
        #      stream is at hand; the two command lines run at the very end, when this process has let go of
        #      its GPU and page-locked memory (a second process with 100 GB on the same GPU slows every driver
        #      call of the command line: context 0.2 -> 1.6 s, teardown 8 ms -> 1 s)
        cli_job = None
        if world == 1 and args.e2e_cli_reads > 0:
            n_c = min(args.e2e_cli_reads, args.reads)
            big = os.path.join(tmp, "e2e_cli.fq")
            t0 = time.perf_counter()
            write_fastq_fixed(big, stream[:n_c * rec].cpu().numpy(), n_c)
            log(f"e2e_cli: wrote {os.path.getsize(big) / 1e9:.2f} GB FASTQ in {time.perf_counter() - t0:.1f} s")
            cli_job = (n_c, big, exe, kind, threads, ncpu)

    # ---- BASELINE config 4: k sweep on reads with N runs (N = 1 only)
    k_sweep = None
    if rank == 0 and world == 1 and args.sweep_reads > 0:
        del stream
        torch.cuda.empty_cache()
        k_sweep = run_k_sweep(torch, vafgpu, dev, ts, args.sweep_reads, peak)
    # ---- BASELINE config 5: the counting mode on a fixed slice (N = 1 only)
    modes = None
    if rank == 0 and world == 1 and args.kc_reads > 0:
        eng.close()
        eng = None
        torch.cuda.empty_cache()
        modes = {"kc": run_kc_mode(args.kc_reads)}

    if rank == 0 and cli_job is not None:
        if eng is not None:
            eng.close()
            eng = None
        host = stream = scratch = None  # the page-locked buffer and the resident stream go too
        torch.cuda.empty_cache()
        n_c, big, exe, kind, threads, ncpu = cli_job
        cli_bases = n_c * READ_LEN
        t_ref = time_reference(exe, pattern_file, big, os.path.join(tmp, "big_ref.vaf"), threads)
        cli_exe = os.path.join(PKG, "vaf-counter")
        ours = {}
        for th in sorted({1, ncpu}):
            out_vaf = os.path.join(tmp, "big_cli%d.vaf" % th)
            t0 = time.perf_counter()
            r = subprocess.run([cli_exe, "-k", str(K), "-t", str(th), "-v", "-p", pattern_file, "-o", out_vaf, big],
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                               env=dict(os.environ, VAFGPU_TIMING="1",  # start-up breakdown on stderr
                                        CUDA_VISIBLE_DEVICES=(os.environ.get("CUDA_VISIBLE_DEVICES") or "0").split(",")[0]))  # N = 1
            wall = time.perf_counter() - t0
            assert open(out_vaf, "rb").read() == open(os.path.join(tmp, "big_ref.vaf"), "rb").read(), \
                "CLI output differs from the reference's at -t %d" % th
            m = re.search(rb"K-mer counting:\s+([0-9.]+) sec", r.stderr)
            ctx_ms = sum(float(x) for x in re.findall(rb"\[vafgpu\] (?:CUDA init \(device count\)|context|module load \+ policy kernel)\s+([0-9.]+) ms", r.stderr))
            ours["t%d" % th] = {"whole_process_s": wall, "whole_process_gbases_s": cli_bases / wall / 1e9,
                                "counting_phase_gbases_s": cli_bases / max(float(m.group(1)), 1e-9) / 1e9 if m else None,
                                "cuda_init_context_module_s": ctx_ms / 1e3,
                                "phases_ms": {a.decode().strip(): float(b) for a, b in re.findall(rb"\[vafgpu\] (.+?)\s+([0-9.]+) ms", r.stderr)},
                                "main_total_s": (lambda t: float(t.group(1)) if t else None)(re.search(rb"Total runtime:\s+([0-9.]+) sec", r.stderr))}
        best_th = min(ours, key=lambda x: ours[x]["whole_process_s"])
        e2e_cli = {"workload": "%d reads x %d bp of this workload as one plain FASTQ file in %s (%.2f GB)"
                               % (n_c, READ_LEN, os.path.dirname(big), os.path.getsize(big) / 1e9),
                   "value": ours[best_th]["whole_process_gbases_s"], "unit": "Gbases/s",
                   "reference": {"value": cli_bases / t_ref / 1e9, "unit": "Gbases/s", "threads": threads, "whole_process_s": t_ref,
                                 "kind": kind},
                   "speedup_whole_process": t_ref / ours[best_th]["whole_process_s"], "this_repo": ours,
                   "vaf_bytes": "identical",
                   "note": "process start (CUDA context creation, 0.5-3 s on these boxes without a persistence daemon: "
                           "this_repo.*.cuda_init_context_module_s), pattern load, FASTQ parse, count and VAF write on both sides"}
        os.unlink(big)
        parity["cli_vaf_bytes_vs_%s_on_%d_reads" % (kind, n_c)] = "identical at -t 1 and -t %d" % ncpu

    if rank == 0:
        traffic, traffic_note = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if prof.get("library_version") == vafgpu.load_library().vafgpu_version().decode():
                traffic = prof["dram_bytes_per_stream_byte"] * (n16 / launches_per_step)
                traffic_note = prof.get("source")
            else:
                traffic_note = "profiles/traffic.json was measured on %r, this library is %r: not used" % (
                    prof.get("library_version"), vafgpu.load_library().vafgpu_version().decode())
        except (OSError, KeyError, ValueError):
            pass
        algo_bytes_per_launch = bases_per_step / launches_per_step            # 1 byte per base
        achieved = algo_bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "Gbases/s", "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 in, u32/u64 integer arithmetic", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": 8 * n_pat,
                    "sample": "%d of %d reads per GPU, already parsed (the path's boundary: no FASTQ parsing, no process "
                              "start; see e2e_cli for those), from page-locked host memory%s through vafgpu_submit_stream + "
                              "vafgpu_finish" % (e2e_reads, args.reads, " on the GPU's NUMA node" if numa is not None else "")},
            "e2e_cli": e2e_cli,
            "strong_scaling": strong,
            "k_sweep": k_sweep,
            "modes": modes,
            "gpu_launches": int(args.steps * launches_per_step),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "kernel": "anchor_scan_kernel<S=%d,CANON=%d,DEFER=%d,L=%d> (%d threads)" % (
                             st_kernel["anchor_stride"], st_kernel["filter_canon"], st_kernel["lookup_deferred"],
                             st_kernel["anchor_len"], st_kernel["kernel_threads"]),
                         "algorithmic_bytes_per_launch": algo_bytes_per_launch, "ms_per_launch": kernel_ms,
                         "ms_per_launch_source": kernel_ms_source},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
            "parity": parity,
            "merge": merge,
            "numa_node": numa,
            "stats": {"hits_per_step": hits_per_step, "resolver_entries_per_base":
                      st_kernel["n_candidates"] / max(st_kernel["n_bytes"], 1), "anchor_stride": st_kernel["anchor_stride"],
                      "anchor_len": st_kernel["anchor_len"], "filter_bytes": st_kernel["filter_bytes"],
                      "filter2_bytes": st_kernel["filter2_bytes"], "pattern_collisions": n_coll,
                      "library": vafgpu.load_library().vafgpu_version().decode()},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if eng is not None:
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
