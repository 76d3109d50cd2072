"""Two linked counting contexts in ONE process, one per GPU, each scanning its own resident stream: the
push form (k-mers appended to their owner's inbox, over NVLink for the peer), for ncu -- which must not
wrap a multi-rank job -- to read the NVLink byte counters of the push kernel.
    ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum -k regex:kc_push python tools/kc_link_prof.py [reads per GPU] [log2 slots per GPU]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmer-cnt_b200"))
import torch

import bench as vb
import kcgpu

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n_dev = min(torch.cuda.device_count(), int(os.environ.get("KC_DEVICES", 2)))
glen = 200_000_000
streams = []
for i in range(n_dev):
    dev = torch.device("cuda", i)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (glen,), device=dev, generator=g)]
    s, _ = vb.make_stream(torch, torch.cat([genome, genome]), glen, n_reads, 5 + i, dev, sub_rate=0.01, n_rate=0.005)
    streams.append(s)
    torch.cuda.synchronize(dev)
ctrs = [kcgpu.Counter(31, 1 << bits, device=i) for i in range(n_dev)]
if n_dev > 1:
    kcgpu.link(ctrs)
t0 = time.perf_counter()
for c, s in zip(ctrs, streams):
    c.count_device(s.data_ptr(), s.numel())
for c in ctrs:
    c.flush()
ms = (time.perf_counter() - t0) * 1e3
tot = None
for c in ctrs:
    h, st = c.histogram()
    tot = h if tot is None else tot + h
    print(f"device {c.device if hasattr(c, 'device') else '?'}: k-mers extracted {st['n_kmers']}, entries made {st['n_distinct']}, direct {st['n_direct']}, flushes {st['n_flushes']}")
print(f"{n_dev} devices x {n_reads} reads: {ms:.1f} ms, {n_dev * n_reads * 150 / ms / 1e6:.1f} Gbases/s, distinct {int(tot.sum())}")
for c in ctrs:
    c.close()
