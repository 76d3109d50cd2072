"""Quick kernel-only timing on a device-resident synthetic stream (development aid)."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import util
from util import vafgpu

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=21)
ap.add_argument("--snps", type=int, default=20920)
ap.add_argument("--mb", type=int, default=2048)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--recipe", action="store_true")
a = ap.parse_args()

rng = np.random.default_rng(1)
pats = util.make_patterns(rng, a.snps, a.k)
keys, vals, _ = vafgpu.build_key_list(pats, a.k)
n = a.mb << 20
n -= n % (151 * 16)
g = torch.Generator(device="cuda"); g.manual_seed(1)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
s = lut[torch.randint(0, 4, (n,), device="cuda", generator=g)]
s[150::151] = 10
s[torch.rand(n, device="cuda", generator=g) < 0.005] = ord("N")
eng = vafgpu.Engine(a.k, keys, vals, a.snps, n_devices=1, flags=vafgpu.F_REFERENCE_RECIPE if a.recipe else 0)
ts_ = torch.cuda.Stream()
torch.cuda.set_stream(ts_)
st = ts_.cuda_stream
assert st != 0
for _ in range(2):
    eng.count_device(s.data_ptr(), n, stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(a.iters):
    e0.record(); eng.count_device(s.data_ptr(), n, stream=st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
counts, stats = eng.finish()
ms = min(ts)
print(f"k={a.k} snps={a.snps} bytes={n} best {ms:.3f} ms  {n/ms/1e6:.1f} GB/s  ({n/ms/1e6/6544.3*100:.1f}% of 6544 GB/s)  all={['%.3f'%t for t in ts]}")
print({k: v for k, v in stats.items()})
