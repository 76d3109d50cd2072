"""One launch of the counting kernel over a small resident stream, for ncu (development aid).
    ncu --set full -k regex:kc_scan -c 1 -o out python tools/kc_prof.py [reads] [log2 slots] [genome] [sub_rate] [direct]
The table must hold the distinct k-mers: about genome + reads * 150 * sub_rate * 31 of them."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmer-cnt_b200"))
import torch

import bench as vb
import kcgpu

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 27
glen = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000_000
sub = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (glen,), device=dev, generator=g)]
stream, _ = vb.make_stream(torch, torch.cat([genome, genome]), glen, n_reads, 5, dev, sub_rate=sub, n_rate=0.005 if sub else 0.0)
torch.cuda.synchronize()
direct = len(sys.argv) > 5 and sys.argv[5] == "direct"
with kcgpu.Counter(31, 1 << bits, list_slots=kcgpu.NO_LISTS if direct else int(os.environ.get("KC_LIST_SLOTS", 0))) as c:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        e0.record()
        c.count_device(stream.data_ptr(), stream.numel(), s.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c.flush()
    flush_ms = (time.perf_counter() - t0) * 1e3
    h, st = c.histogram()
    ms = e0.elapsed_time(e1)
    tot = ms + flush_ms
    print(f"{n_reads} reads, 2^{bits} slots{' direct' if direct else ''}: scan {ms:.3f} ms + flush {flush_ms:.3f} ms, "
          f"{st['n_kmers'] / tot / 1e6:.2f} G k-mers/s, {n_reads * 150 / tot / 1e6:.2f} Gbases/s, distinct {st['n_distinct']}, "
          f"load {st['n_distinct'] / (1 << bits):.3f}, direct {st['n_direct']}, flushes {st['n_flushes']}, "
          f"histogram crc {__import__('zlib').crc32(h.tobytes()):08x} ({os.path.basename(os.environ.get('VAFGPU_LIB', 'libvafgpu.so'))})")
