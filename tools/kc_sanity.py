"""A small run of every counting kernel -- one owner (tile scan, list insert), three linked owners on one device (push,
route, list insert), the list-less form, the filtered two-pass form -- checked against the oracle; small enough for
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/kc_sanity.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
from util import kcgpu

rng = np.random.default_rng(5)
reads = util.make_genome_reads(rng, 30000, 4000, jitter=60, junk_rate=0.003, lower_rate=0.05, n_rate=0.01, repeat=6)
k = 31
want, n_inst, n_dist = util.KcOracle().count_reads(reads, k)


def fill(ctrs):
    for i, r in enumerate(reads):
        ctrs[i % len(ctrs)].add_read(r)


for name, make in (("one owner", lambda: [kcgpu.Counter(k, 1 << 20, block_bytes=1 << 16)]),
                   ("one owner, short lists", lambda: [kcgpu.Counter(k, 1 << 20, block_bytes=1 << 16, list_slots=1 << 12)]),
                   ("three owners", lambda: [kcgpu.Counter(k, 1 << 19, block_bytes=1 << 16) for _ in range(3)]),
                   ("no lists", lambda: [kcgpu.Counter(k, 1 << 20, block_bytes=1 << 16, list_slots=kcgpu.NO_LISTS)])):
    ctrs = make()
    if len(ctrs) > 1:
        kcgpu.link(ctrs)
    fill(ctrs)
    tot, inst = np.zeros(256, dtype=np.uint64), 0
    for c in ctrs:
        h, st = c.histogram()
        tot += h
        inst += st["n_kmers"]
    for c in ctrs:
        c.close()
    assert np.array_equal(tot, want) and inst == n_inst, name
    print(f"{name}: {inst} k-mers, histogram matches the oracle")
want_yak = util.YakOracle().count_reads(reads, k, bf_shift=24)
with kcgpu.Counter(k, 1 << 20, block_bytes=1 << 16, bloom_bits=22, bloom_hashes=4) as c:
    c.set_pass(kcgpu.PASS_CLAIM)
    fill([c])
    c.set_pass(kcgpu.PASS_LOOKUP)
    fill([c])
    got, _ = c.histogram1024(2, 1023)
assert np.array_equal(got, want_yak)
print("filtered two-pass count matches the oracle")
