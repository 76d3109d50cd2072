"""The library's own several-GPU paths on a box with >= 2 B200s (run with `gpurun --gpus 2`;
skipped on one GPU): blocks dealt round-robin inside one process with every kernel adding into
device 0's counter vector over NVLink (vafgpu_create(n_devices > 1)), and one process per GPU
sharing one vector through CUDA IPC (vafgpu_export_counters / vafgpu_attach_counters).  Both
stand in for the all-reduce of per-device vectors; uint32 sums commute, so the counts must be
bit-identical to the one-GPU result and to the oracle (vaf-counter.c:449-479)."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from util import vafgpu

pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")


def cfg2_case(tmp_path, oracle, n_reads, seed):
    pf = str(tmp_path / "cfg2_patterns.txt")
    with gzip.open(os.path.join(util.GOLDEN, "cfg2_patterns.txt.gz"), "rb") as src, open(pf, "wb") as dst:
        dst.write(src.read())
    pats = vafgpu.load_patterns(pf)
    rng = np.random.default_rng(seed)
    reads = util.make_reads(rng, pats, 21, n_reads, mean_len=150, jitter=20, plant=0.9, n_rate=0.005)
    want, _, _ = oracle.count_reads(pf, 21, reads)
    keys, vals, _ = vafgpu.build_key_list(pats, 21)
    return pf, pats, reads, want, keys, vals


@needs2
@pytest.mark.parametrize("flags", [0, vafgpu.F_HOST_MERGE])
def test_round_robin_over_devices_one_vector(tmp_path, oracle, lib, flags):
    pf, pats, reads, want, keys, vals = cfg2_case(tmp_path, oracle, 40000, 5)
    nd = n_gpus()
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=nd, block_bytes=1 << 16, flags=flags) as eng:
        half = len(reads) // 2
        for r in reads[:half]:
            eng.add_read(r)
        first, _ = eng.finish()
        for r in reads[half:]:
            eng.add_read(r)
        total, st = eng.finish()
        eng.reset()
        zero, _ = eng.finish()
    assert st["n_devices"] == nd and st["n_blocks"] >= 80 and st["lookup_deferred"] == 1
    w1, _, _ = oracle.count_reads(pf, 21, reads[:half])
    assert np.array_equal(first, w1)
    assert np.array_equal(total, want)
    assert not zero.any()


@needs2
def test_submit_stream_deals_blocks_to_every_device(tmp_path, oracle, lib):
    torch = pytest.importorskip("torch")
    pf, pats, reads, want, keys, vals = cfg2_case(tmp_path, oracle, 30000, 6)
    stream = util.pack_stream(reads, 21)
    host = torch.from_numpy(stream).pin_memory()
    nd = n_gpus()
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=nd, block_bytes=1 << 17) as eng:
        eng.submit_stream((host.data_ptr(), host.numel()), n_reads=len(reads), n_bases=sum(map(len, reads)))
        got, st = eng.finish()
    assert np.array_equal(got, want)
    assert st["n_blocks"] >= 2 * nd


CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import util
from util import vafgpu
pf, handle, fq = sys.argv[3], bytes.fromhex(sys.argv[4]), sys.argv[5]
pats = vafgpu.load_patterns(pf)
keys, vals, _ = vafgpu.build_key_list(pats, 21)
reads = open(fq, "rb").read().split(b"\n")
with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1, block_bytes=1 << 16) as eng:
    eng.attach_counters(handle)
    for r in reads:
        eng.add_read(r)
    own, st = eng.finish()
    assert not own.any()
    print("child blocks", st["n_blocks"])
"""


@needs2
def test_one_process_per_gpu_shares_one_vector_over_ipc(tmp_path, oracle, lib):
    """rank 1 (its own process, its own GPU) adds into rank 0's vector; rank 0 reads the total"""
    pf, pats, reads, want, keys, vals = cfg2_case(tmp_path, oracle, 30000, 8)
    half = len(reads) // 2
    fq = str(tmp_path / "second_half.txt")
    with open(fq, "wb") as fh:
        fh.write(b"\n".join(reads[half:]))
    with vafgpu.Engine(21, keys, vals, len(pats), n_devices=1, block_bytes=1 << 16) as eng:
        handle = eng.export_counters()
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="1")
        child = subprocess.Popen([sys.executable, "-c", CHILD, os.path.join(util.ROOT, "tests"), util.PKG, pf, handle.hex(), fq],
                                 env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        for r in reads[:half]:
            eng.add_read(r)
        out, err = child.communicate(timeout=300)
        assert child.returncode == 0, err.decode()[-2000:]
        got, _ = eng.finish()          # after the child has drained (its exit is the barrier)
    assert np.array_equal(got, want)


@needs2
def test_command_lines_on_two_gpus_print_what_they_print_on_one(tmp_path, lib):
    """the three command lines use one GPU unless told (or forced by the table size) to take more: with
    VAFGPU_DEVICES / KCGPU_DEVICES = 2 the bytes they write are the ones they write on one GPU"""
    rng = np.random.default_rng(77)
    pats = util.make_patterns(rng, 300, 21)
    reads = util.make_reads(rng, pats, 21, 30000, mean_len=150, jitter=30, junk_rate=0.003)
    pf, fq = str(tmp_path / "p.txt"), str(tmp_path / "r.fq")
    util.write_patterns(pf, pats)
    with open(fq, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)))
    outs = {}
    for nd in ("1", "2"):
        out = str(tmp_path / ("o%s.vaf" % nd))
        r = subprocess.run([os.path.join(util.PKG, "vaf-counter"), "-k", "21", "-t", "4", "-v", "-p", pf, "-o", out, fq], check=True,
                           capture_output=True, env=dict(os.environ, VAFGPU_DEVICES=nd))
        assert ("Devices:               %s" % nd).encode() in r.stderr
        outs["vaf" + nd] = open(out, "rb").read()
        for exe, args in (("kc-c4", ["-k", "31", "-t", "4"]), ("yak-count", ["-k", "31", "-t", "4", "-b", "26"])):
            outs[exe + nd] = subprocess.run([os.path.join(util.PKG, exe)] + args + [fq], check=True, capture_output=True,
                                            env=dict(os.environ, KCGPU_DEVICES=nd)).stdout
    for name in ("vaf", "kc-c4", "yak-count"):
        assert outs[name + "1"] == outs[name + "2"] and len(outs[name + "1"]) > 100, name
