"""Host-side logic that needs no GPU: the C ABI loads and exports what include/vafgpu.h
declares, refuses to run without a device (no CPU fallback), the read canonicaliser follows
the reference's byte rules, and the table builder + anchor bookkeeping of the CUDA kernel,
emulated on the host by tests/cpu_sim (test-only), reproduce the oracle's counts."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import util
from util import vafgpu


def test_library_exports_every_symbol_of_the_header(lib):
    hdr = open(os.path.join(util.ROOT, "include", "vafgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(vafgpu_[a-z_]+)\s*\(", hdr)))
    assert set(declared) == set(vafgpu.EXPORTS), declared
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", vafgpu.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name
    assert b"sm_100a" in lib.vafgpu_version()


def test_library_carries_sm100a_sass_with_the_expected_mnemonics(lib):
    sass = subprocess.run(["cuobjdump", "-sass", vafgpu.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "anchor_scan_kernel" in sass and "recipe_scan_kernel" in sass
    assert "LDG.E.EF.128" in sass          # 128-bit streaming loads
    assert "CCTL.E.PF2" in sass            # prefetch.global.L2
    assert "MATCH.ANY" in sass             # warp-aggregated atomics


def test_no_device_means_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    keys = np.array([0x11d011da169], dtype=np.uint64)
    vals = np.array([0], dtype=np.uint32)
    with pytest.raises(vafgpu.VafGpuError) as ei:
        vafgpu.Engine(21, keys, vals, 1)
    assert ei.value.code == vafgpu.ENOGPU


def test_argument_validation(lib):
    h = C.c_void_p()
    k64 = (C.c_uint64 * 1)(1 << 50)
    v32 = (C.c_uint32 * 1)(0)
    assert lib.vafgpu_create(C.byref(h), 0, k64, v32, 1, 1, 0, 0, 0, 0) == vafgpu.EINVAL
    assert lib.vafgpu_create(C.byref(h), 32, k64, v32, 1, 1, 0, 0, 0, 0) == vafgpu.EINVAL
    assert lib.vafgpu_create(C.byref(h), 21, k64, v32, 1, 1, 0, 0, 0, 0) == vafgpu.EINVAL   # key wider than 2k bits
    v32[0] = 2
    k64[0] = 5
    assert lib.vafgpu_create(C.byref(h), 21, k64, v32, 1, 1, 0, 0, 0, 0) == vafgpu.EINVAL   # value names pattern 1 of 1
    assert b"pattern" in lib.vafgpu_strerror(None)
    assert lib.vafgpu_add_read(None, b"ACGT", 4) == vafgpu.EINVAL


def test_anchor_plan():
    assert vafgpu.plan(21) == (8, 14)
    assert vafgpu.plan(15) == (4, 12)
    assert vafgpu.plan(31) == (16, 16)
    for k in range(1, 32):
        s, l = vafgpu.plan(k)
        assert s in (1, 2, 4, 8, 16) and 1 <= l <= 16 and l <= k - s + 1   # the anchor lies inside every alignment
    with pytest.raises(vafgpu.VafGpuError):
        vafgpu.plan(32)


def test_canonicaliser_follows_the_reference_byte_rules(oracle):
    """vafgpu_canonicalise_read must turn a read into A/C/G/T/N such that decoding the result
    with the strict table gives exactly the codes the reference's encoder gives the original
    (low-nibble rule below len & ~15, strict table for the tail; vaf-counter.c:261-291)."""
    rng = np.random.default_rng(0)
    allbytes = bytes(range(256))
    reads = [allbytes, allbytes[::-1], allbytes[:250], allbytes[:17], allbytes[:16], allbytes[:15], b""]
    reads += [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in rng.integers(1, 80, 200)]
    for r in reads:
        for simd in (True, False):
            out = vafgpu.canonicalise_read(r, simd)
            assert len(out) == len(r) and set(out) <= set(b"ACGTN")
            for i, (b, o) in enumerate(zip(r, out)):
                want = (oracle.lib.vo_nt4_nibble(b) if simd and i < (len(r) & ~15) else oracle.lib.vo_nt4_strict(b))
                assert "ACGTN"[want] == chr(o), (i, b, o, simd)


def test_pack16_and_revcomp_arithmetic(sim):
    rng = np.random.default_rng(3)
    code = {ord("A"): 0, ord("C"): 1, ord("T"): 2, ord("G"): 3, ord("a"): 0, ord("c"): 1, ord("t"): 2, ord("g"): 3, ord("U"): 2}
    alphabet = np.frombuffer(b"ACGTacgtU", dtype=np.uint8)
    for _ in range(2000):
        chunk = alphabet[rng.integers(0, len(alphabet), 16)].tobytes()
        want = sum(code[b] << (2 * i) for i, b in enumerate(chunk))
        assert sim.lib.sim_pack16(chunk) == want
    for L in range(1, 17):
        for _ in range(200):
            x = int(rng.integers(0, 1 << (2 * L)))
            bases = [(x >> (2 * i)) & 3 for i in range(L)]
            rc = sum(((b ^ 2) << (2 * (L - 1 - i))) for i, b in enumerate(bases))
            assert sim.lib.sim_rc32(x, L) == rc
            assert sim.lib.sim_rc32(rc, L) == x


@pytest.mark.parametrize("k", [1, 2, 5, 11, 12, 13, 14, 15, 16, 18, 19, 21, 22, 23, 26, 27, 29, 31])
def test_anchor_algorithm_emulation_matches_oracle(tmp_path, oracle, sim, k):
    rng = np.random.default_rng(k)
    n = 300
    pats = util.make_patterns(rng, n, k, dup_every=37, bad_every=41)
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    reads = util.make_reads(rng, pats, k, 2500, mean_len=100, jitter=60, junk_rate=0.01, lower_rate=0.02)
    want, _, ncoll = oracle.count_reads(pf, k, reads)
    keys, vals, ncoll2 = vafgpu.build_key_list(vafgpu.load_patterns(pf), k)
    assert ncoll == ncoll2
    got, n_cand, info = sim.count(k, keys, vals, n, util.pack_stream(reads, k))
    assert np.array_equal(got, want)
    assert want.sum() > 0 and n_cand >= want.sum()
    assert (info[0], info[1]) == vafgpu.plan(k)


def test_large_panel_switches_to_strand_symmetric_filter_keys(tmp_path, oracle, sim):
    rng = np.random.default_rng(9)
    pats = util.make_patterns(rng, 21000, 21)
    pf = str(tmp_path / "p.txt")
    util.write_patterns(pf, pats)
    reads = util.make_reads(rng, pats, 21, 3000, plant=0.8)
    want, _, _ = oracle.count_reads(pf, 21, reads)
    keys, vals, _ = vafgpu.build_key_list(vafgpu.load_patterns(pf), 21)
    got, n_cand, info = sim.count(21, keys, vals, len(pats), util.pack_stream(reads, 21))
    assert np.array_equal(got, want)
    assert info[5] & 0x80000000                      # canon filter
    assert (info[5] & 0x7fffffff) <= 8 * 2 * 21000   # one key serves both strands
    assert info[2] * 4 <= 227 * 1024                 # fits the shared-memory budget


def test_pattern_loader_and_vaf_writer_mirror_the_reference(tmp_path, oracle):
    d = os.path.join(util.GOLDEN, "e2e_k21")
    pats = vafgpu.load_patterns(os.path.join(d, "patterns.txt"))
    lines = open(os.path.join(d, "expected.vaf")).read().splitlines()
    assert len(pats) == len(lines) - 2
    counts = np.zeros(2 * len(pats), dtype=np.uint32)
    for i, l in enumerate(lines[2:]):
        f = l.split("\t")
        counts[2 * i], counts[2 * i + 1] = int(f[5]), int(f[6])
    assert vafgpu.format_vaf(pats, counts) == open(os.path.join(d, "expected.vaf")).read()
    # truncated file: the first malformed record ends the load
    p = tmp_path / "bad.txt"
    p.write_text(open(os.path.join(d, "patterns.txt")).read().split("\n", 3)[0] + "\nchr1\tnotanumber\t3\trs\tA\tC\tAAA\tCCC\n")
    assert len(vafgpu.load_patterns(str(p))) == 1
    # uint32 wrap-around of the total column and the 0.0000 VAF of an empty row (vaf-counter.c:673-674)
    txt = vafgpu.format_vaf(pats[:2], np.array([0xFFFFFFFF, 2, 0, 0], dtype=np.uint32))
    assert txt.splitlines()[2].split("\t")[5:] == ["4294967295", "2", "1", "2.0000"]
    assert txt.splitlines()[3].endswith("\t0\t0\t0\t0.0000")
