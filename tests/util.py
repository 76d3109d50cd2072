"""Shared helpers for the test-suite: ctypes bindings of the oracle (test infrastructure) and of
the test-only host emulation, seeded synthetic cases, stream packing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kmer-cnt_b200")
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
BUILD_DIR = os.path.join(ROOT, "tests", "_build")
GOLDEN = os.path.join(ROOT, "tests", "golden")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import vafgpu  # noqa: E402


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle.so", "vaf_oracle", "kc_oracle", "yak_oracle", "spg_oracle", "synth"],
                   check=True)


def build_sim() -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    out = os.path.join(BUILD_DIR, "libanchor_sim.so")
    srcs = [os.path.join(ROOT, "tests", "cpu_sim", "anchor_sim.cpp"),
            os.path.join(PKG, "csrc", "vafgpu_tables.cpp")]
    deps = srcs + [os.path.join(PKG, "csrc", "vafgpu_common.h"), os.path.join(PKG, "csrc", "vafgpu_tables.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out] + srcs, check=True)
    return out


def build_ingest_stub() -> str:
    """tests/_build/libingest_stub.so: host/ingest.c + host/fastx.c over a stand-in engine (test-only)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    out = os.path.join(BUILD_DIR, "libingest_stub.so")
    srcs = [os.path.join(ROOT, "tests", "cpu_sim", "ingest_stub.c"), os.path.join(PKG, "host", "ingest.c"),
            os.path.join(PKG, "host", "fastx.c")]
    deps = srcs + [os.path.join(PKG, "host", "ingest.h"), os.path.join(PKG, "host", "fastx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", out] + srcs + ["-lz", "-lpthread"], check=True)
    return out


def stub_ingest(files, k, block_len, n_threads, slice_bytes=None):
    """Run host/ingest.c over `files`; returns (digest tuple, per-file seqs, per-file bases, per-file slices)."""
    lib = C.CDLL(build_ingest_stub())
    lib.stub_ingest.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    n = len(files)
    arr = (C.c_char_p * n)(*[f.encode() for f in files])
    out = (C.c_uint64 * (4 + 3 * n))()
    old = os.environ.get("VAFGPU_SLICE_BYTES")
    if slice_bytes:
        os.environ["VAFGPU_SLICE_BYTES"] = str(slice_bytes)
    try:
        rc = lib.stub_ingest(n, arr, k, block_len, n_threads, out)
    finally:
        if slice_bytes:
            if old is None:
                del os.environ["VAFGPU_SLICE_BYTES"]
            else:
                os.environ["VAFGPU_SLICE_BYTES"] = old
    assert rc == 0
    o = list(out)
    return tuple(o[:4]), o[4:4 + n], o[4 + n:4 + 2 * n], [x & 0xFFFFFFFF for x in o[4 + 2 * n:]]


class VoPattern(C.Structure):
    _fields_ = [("chr", C.c_char * 256), ("start", C.c_int), ("end", C.c_int), ("rsid", C.c_char * 256),
                ("ref", C.c_char), ("alt", C.c_char), ("ref_kmer", C.c_char * 128), ("alt_kmer", C.c_char * 128),
                ("ref_count", C.c_uint32), ("alt_count", C.c_uint32)]


class VoPatterns(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("a", C.POINTER(VoPattern))]


class VoMap(C.Structure):
    _fields_ = [("bits", C.c_uint32), ("count", C.c_uint32), ("used", C.POINTER(C.c_uint32)),
                ("key", C.POINTER(C.c_uint64)), ("val", C.POINTER(C.c_uint32)), ("n_collisions", C.c_int)]


class VoStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_kmers", C.c_uint64)]


class Oracle:
    """liboracle.so: the CPU restatement of the reference path (oracle/vaf_oracle.h)."""

    def __init__(self):
        build_oracle()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        lib.vo_nt4_strict.argtypes = [C.c_uint8]
        lib.vo_nt4_nibble.argtypes = [C.c_uint8]
        lib.vo_encode_kmer.argtypes = [C.c_char_p, C.c_int]
        lib.vo_encode_kmer.restype = C.c_uint64
        lib.vo_revcomp.argtypes = [C.c_uint64, C.c_int]
        lib.vo_revcomp.restype = C.c_uint64
        lib.vo_canonical.argtypes = [C.c_uint64, C.c_int]
        lib.vo_canonical.restype = C.c_uint64
        lib.vo_kmer_hash.argtypes = [C.c_uint64]
        lib.vo_kmer_hash.restype = C.c_uint32
        lib.vo_h2b.argtypes = [C.c_uint32, C.c_uint32]
        lib.vo_h2b.restype = C.c_uint32
        lib.vo_load_patterns.argtypes = [C.c_char_p]
        lib.vo_load_patterns.restype = C.POINTER(VoPatterns)
        lib.vo_patterns_free.argtypes = [C.POINTER(VoPatterns)]
        lib.vo_map_build.argtypes = [C.POINTER(VoPatterns), C.c_int]
        lib.vo_map_build.restype = C.POINTER(VoMap)
        lib.vo_map_free.argtypes = [C.POINTER(VoMap)]
        lib.vo_map_get.argtypes = [C.POINTER(VoMap), C.c_uint64]
        lib.vo_map_get.restype = C.c_uint32
        lib.vo_map_export.argtypes = [C.POINTER(VoMap), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        lib.vo_map_export.restype = C.c_uint32
        lib.vo_count_read.argtypes = [C.POINTER(VoMap), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        lib.vo_count_read.restype = C.c_uint64
        lib.vo_extract_read.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        lib.vo_extract_read.restype = C.c_uint64
        lib.vo_count_file.argtypes = [C.POINTER(VoMap), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint32), C.POINTER(VoStats)]
        lib.vo_count_file.restype = C.c_int
        self.lib = lib

    def extract(self, seq: bytes, k: int, simd: bool = True) -> List[int]:
        out = (C.c_uint64 * max(len(seq), 1))()
        n = self.lib.vo_extract_read(k, seq, len(seq), int(simd), out)
        return [int(out[i]) for i in range(n)]

    def count_reads(self, pattern_file: str, k: int, reads: Sequence[bytes], simd: bool = True) -> Tuple[np.ndarray, int, int]:
        """(counts[2n], n_kmers, n_collisions) of the reference recipe over `reads` (len >= k only)."""
        db = self.lib.vo_load_patterns(pattern_file.encode())
        assert db, pattern_file
        m = self.lib.vo_map_build(db, k)
        n = db.contents.n
        counts = np.zeros(2 * max(n, 1), dtype=np.uint32)
        cp = counts.ctypes.data_as(C.POINTER(C.c_uint32))
        nk = 0
        for r in reads:
            if len(r) >= k:
                nk += self.lib.vo_count_read(m, k, r, len(r), int(simd), cp)
        ncoll = m.contents.n_collisions
        self.lib.vo_map_free(m)
        self.lib.vo_patterns_free(db)
        return counts[: 2 * n], int(nk), int(ncoll)

    def count_file(self, pattern_file: str, k: int, fastx: str, *, simd: bool = True, threads: int = 1,
                   block_len: int = 10_000_000):
        db = self.lib.vo_load_patterns(pattern_file.encode())
        m = self.lib.vo_map_build(db, k)
        n = db.contents.n
        counts = np.zeros(2 * max(n, 1), dtype=np.uint32)
        st = VoStats()
        rc = self.lib.vo_count_file(m, k, fastx.encode(), int(simd), threads, block_len,
                                    counts.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(st))
        self.lib.vo_map_free(m)
        self.lib.vo_patterns_free(db)
        return rc, counts[: 2 * n], {"n_reads": st.n_reads, "n_bases": st.n_bases, "n_kmers": st.n_kmers}


class Sim:
    """tests/_build/libanchor_sim.so: host emulation of the anchor kernel (test-only)."""

    def __init__(self):
        lib = C.CDLL(build_sim())
        lib.sim_anchor_count.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32,
                                         C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.sim_anchor_count.restype = C.c_uint64
        lib.sim_recipe_get.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32,
                                       C.c_uint64, C.POINTER(C.c_uint32)]
        lib.sim_recipe_get.restype = C.c_int
        lib.sim_pack16.argtypes = [C.c_char_p]
        lib.sim_pack16.restype = C.c_uint32
        lib.sim_rc32.argtypes = [C.c_uint32, C.c_int]
        lib.sim_rc32.restype = C.c_uint32
        self.lib = lib

    def count(self, k: int, keys: np.ndarray, vals: np.ndarray, n_patterns: int, stream: np.ndarray):
        counts = np.zeros(2 * max(n_patterns, 1), dtype=np.uint32)
        info = (C.c_uint32 * 6)()
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        ncand = self.lib.sim_anchor_count(k, keys.ctypes.data_as(C.POINTER(C.c_uint64)),
                                          vals.ctypes.data_as(C.POINTER(C.c_uint32)), keys.size,
                                          stream.ctypes.data, stream.size,
                                          counts.ctypes.data_as(C.POINTER(C.c_uint32)), info)
        return counts[: 2 * n_patterns], int(ncand), list(info)


# ---------------------------------------------------------------------------------------------
# synthetic cases

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b
JUNK = np.frombuffer(b"RYKMSWBDHVacgtnUu.-*XQ1357 \x00\x01\x02\x03", dtype=np.uint8)


def revcomp(b: bytes) -> bytes:
    return COMP[np.frombuffer(b, dtype=np.uint8)][::-1].tobytes()


def make_patterns(rng: np.random.Generator, n: int, k: int, *, dup_every: int = 0, bad_every: int = 0):
    """n pattern rows with random k-mers; alt differs from ref at the centre base."""
    pats = []
    for i in range(n):
        ref = ACGT[rng.integers(0, 4, k)].copy()
        alt = ref.copy()
        mid = k // 2
        alt[mid] = ACGT[(int(np.where(ACGT == ref[mid])[0][0]) + int(rng.integers(1, 4))) % 4]
        if dup_every and i and i % dup_every == 0:  # same position under two ids: identical k-mers
            ref = np.frombuffer(pats[-1].ref_kmer.encode(), dtype=np.uint8).copy()
        alt_s = alt.tobytes().decode()
        if bad_every and i % bad_every == 3:
            alt_s = alt_s[:2] + "N" + alt_s[3:]  # unusable k-mer: never inserted
        pats.append(vafgpu.Pattern("chr%d" % (1 + i % 22), 1000 + 7 * i, 1001 + 7 * i, "rs%d" % i,
                                   chr(ref[mid]), chr(alt[mid]), ref.tobytes().decode(), alt_s))
    return pats


def write_patterns(path: str, pats) -> None:
    with open(path, "w") as fh:
        for p in pats:
            fh.write("%s\t%d\t%d\t%s\t%s\t%s\t%s\t%s\n" % (p.chr, p.start, p.end, p.rsid, p.ref, p.alt, p.ref_kmer, p.alt_kmer))


def make_reads(rng: np.random.Generator, pats, k: int, n_reads: int, *, mean_len: int = 150, jitter: int = 0,
               plant: float = 0.5, n_rate: float = 0.005, junk_rate: float = 0.0, lower_rate: float = 0.0) -> List[bytes]:
    """Random reads; a fraction carries a pattern k-mer (either allele, either strand) at a
    random offset, possibly clipped by the read end or broken by an N."""
    reads = []
    for _ in range(n_reads):
        ln = mean_len + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0)
        ln = max(ln, 0)
        s = ACGT[rng.integers(0, 4, ln)].copy()
        if pats and ln and rng.random() < plant:
            for _ in range(int(rng.integers(1, 3))):
                p = pats[int(rng.integers(0, len(pats)))]
                km = (p.alt_kmer if rng.random() < 0.5 else p.ref_kmer).encode()
                if b"N" in km:
                    continue
                if rng.random() < 0.5:
                    km = revcomp(km)
                at = int(rng.integers(-k // 2, ln))
                lo, hi = max(at, 0), min(at + k, ln)
                if hi > lo:
                    s[lo:hi] = np.frombuffer(km, dtype=np.uint8)[lo - at:hi - at]
        if n_rate:
            s[rng.random(ln) < n_rate] = ord("N")
        if junk_rate:
            m = rng.random(ln) < junk_rate
            s[m] = JUNK[rng.integers(0, len(JUNK), int(m.sum()))]
        if lower_rate:
            m = rng.random(ln) < lower_rate
            s[m] = s[m] | 0x20
        reads.append(s.tobytes())
    return reads


def pack_stream(reads: Sequence[bytes], k: int, simd_rule: bool = True) -> np.ndarray:
    """What vafgpu_add_read builds in a staging block: canonicalised reads, '\\n' separated,
    padded with '\\n' to a multiple of 16."""
    parts = []
    for r in reads:
        if len(r) >= k:
            parts.append(vafgpu.canonicalise_read(r, simd_rule))
            parts.append(b"\n")
    buf = b"".join(parts)
    buf += b"\n" * (-len(buf) % 16)
    return np.frombuffer(buf, dtype=np.uint8).copy()


def write_fastq(path: str, reads: Sequence[bytes], *, fasta: bool = False, line: int = 0) -> None:
    with open(path, "wb") as fh:
        for i, r in enumerate(reads):
            if fasta:
                fh.write(b">r%d\n" % i)
                if line:
                    for j in range(0, len(r), line):
                        fh.write(r[j:j + line] + b"\n")
                else:
                    fh.write(r + b"\n")
            else:
                fh.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)))


# ---------------------------------------------------------------------------------------------
# counting mode (kc-c4 path)

import kcgpu  # noqa: E402


class KcOracle:
    """liboracle.so's restatement of kc-c4 (oracle/kc_oracle.h)."""

    def __init__(self):
        build_oracle()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        lib.kco_nt4.argtypes = [C.c_uint8]
        lib.kco_hash64.argtypes = [C.c_uint64, C.c_int]
        lib.kco_hash64.restype = C.c_uint64
        lib.kco_hashed_kmers.argtypes = [C.c_char_p, C.c_long, C.c_int, C.POINTER(C.c_uint64)]
        lib.kco_hashed_kmers.restype = C.c_long
        lib.kco_create.argtypes = [C.c_int]
        lib.kco_create.restype = C.c_void_p
        lib.kco_destroy.argtypes = [C.c_void_p]
        lib.kco_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        lib.kco_add_hashed.argtypes = [C.c_void_p, C.c_uint64]
        lib.kco_add_file.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        lib.kco_hist.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.kco_distinct.argtypes = [C.c_void_p]
        lib.kco_distinct.restype = C.c_uint64
        lib.kco_instances.argtypes = [C.c_void_p]
        lib.kco_instances.restype = C.c_uint64
        self.lib = lib

    def hashed_kmers(self, seq: bytes, k: int) -> np.ndarray:
        out = np.zeros(max(len(seq), 1), dtype=np.uint64)
        n = self.lib.kco_hashed_kmers(seq, len(seq), k, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out[:n]

    def _finish(self, o) -> Tuple[np.ndarray, int, int]:
        hist = np.zeros(256, dtype=np.uint64)
        self.lib.kco_hist(o, hist.ctypes.data_as(C.POINTER(C.c_uint64)))
        res = hist, int(self.lib.kco_instances(o)), int(self.lib.kco_distinct(o))
        self.lib.kco_destroy(o)
        return res

    def count_reads(self, reads: Sequence[bytes], k: int) -> Tuple[np.ndarray, int, int]:
        """(hist[256], k-mer instances, distinct k-mers) of kc-c4's recipe over `reads`."""
        o = self.lib.kco_create(k)
        for r in reads:
            self.lib.kco_add_read(o, r, len(r))
        return self._finish(o)

    def count_file(self, fn: str, k: int, block_len: int = 10_000_000) -> Tuple[np.ndarray, int, int]:
        o = self.lib.kco_create(k)
        assert self.lib.kco_add_file(o, fn.encode(), block_len) == 0, fn
        return self._finish(o)

    def count_hashed(self, hashed: np.ndarray, k: int) -> Tuple[np.ndarray, int, int]:
        o = self.lib.kco_create(k)
        for h in np.asarray(hashed, dtype=np.uint64).tolist():
            self.lib.kco_add_hashed(o, h)
        return self._finish(o)


class YakOracle:
    """liboracle.so's restatement of yak-count (oracle/yak_oracle.h)."""

    def __init__(self):
        build_oracle()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        lib.yko_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        lib.yko_create.restype = C.c_void_p
        lib.yko_destroy.argtypes = [C.c_void_p]
        lib.yko_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_int]
        lib.yko_second_pass.argtypes = [C.c_void_p]
        lib.yko_shrink.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.yko_shrink.restype = C.c_uint64
        lib.yko_count_files.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_long]
        lib.yko_hist.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.yko_distinct.argtypes = [C.c_void_p]
        lib.yko_distinct.restype = C.c_uint64
        lib.yko_bf_insert.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_uint64]
        self.lib = lib

    def _hist(self, o) -> np.ndarray:
        hist = np.zeros(1024, dtype=np.uint64)
        self.lib.yko_hist(o, hist.ctypes.data_as(C.POINTER(C.c_uint64)))
        self.lib.yko_destroy(o)
        return hist

    def count_reads(self, reads: Sequence[bytes], k: int, *, bf_shift: int = 0, n_hash: int = 4, pre: int = 10,
                    reads2: Optional[Sequence[bytes]] = None) -> np.ndarray:
        """hist[1024] of yak_count_file (yak-count.c:445-456) over in-memory reads"""
        o = self.lib.yko_create(k, pre, bf_shift, n_hash)
        for r in reads:
            self.lib.yko_add_read(o, r, len(r), 1)
        if bf_shift > 0:
            self.lib.yko_second_pass(o)
            for r in (reads if reads2 is None else reads2):
                self.lib.yko_add_read(o, r, len(r), 0)
            self.lib.yko_shrink(o, 2, 1023)
        return self._hist(o)

    def count_files(self, fn1: str, fn2: Optional[str], k: int, *, bf_shift: int = 0, n_hash: int = 4, pre: int = 10,
                    chunk: int = 10_000_000) -> np.ndarray:
        o = self.lib.yko_create(k, pre, bf_shift, n_hash)
        assert self.lib.yko_count_files(o, fn1.encode(), fn2.encode() if fn2 else None, chunk) == 0
        return self._hist(o)


def build_kc_sim() -> str:
    """tests/_build/libkc_sim.so: host emulation of the counting mode's bookkeeping (test-only)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    out = os.path.join(BUILD_DIR, "libkc_sim.so")
    src = os.path.join(ROOT, "tests", "cpu_sim", "kc_sim.cpp")
    deps = [src, os.path.join(PKG, "csrc", "kcgpu_kernels.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + cuda_inc, "-o", out, src], check=True)
    return out


class KcSim:
    def __init__(self):
        lib = C.CDLL(build_kc_sim())
        lib.sim_kc_geometry.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
        lib.sim_kc_hash64.argtypes = [C.c_uint64, C.c_int]
        lib.sim_kc_hash64.restype = C.c_uint64
        lib.sim_kc_count.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.sim_kc_count.restype = C.c_uint64
        lib.sim_kc_extract.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.sim_kc_extract.restype = C.c_uint64
        lib.sim_kc_tile_run.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]
        lib.sim_kc_tile_run.restype = C.c_uint64
        self.lib = lib

    def extract(self, k, stream: np.ndarray) -> np.ndarray:
        """hash64 of every canonical k-mer of a packed stream, by the tile kernels' own extraction"""
        assert stream.size % 16 == 0
        out = np.zeros(stream.size, dtype=np.uint64)
        n = self.lib.sim_kc_extract(k, stream.ctypes.data, stream.size, out.ctypes.data, out.size)
        return out[:n]

    def geometry(self, k, n_slots, list_cap, region_bits):
        out = (C.c_uint64 * 7)()
        self.lib.sim_kc_geometry(k, n_slots, list_cap, region_bits, out)
        return dict(zip(("need_bits", "alloc", "lists", "cursors", "inbox_cursor", "inbox_cap", "lists2_end"), map(int, out)))

    def count(self, k, n_parts, table_bits, list_cap, stream: np.ndarray):
        hist = np.zeros(256, dtype=np.uint64)
        nd = C.c_uint64()
        lost = self.lib.sim_kc_count(k, n_parts, table_bits, list_cap, stream.ctypes.data, stream.size,
                                     hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(nd))
        return hist, int(lost), int(nd.value)


def pack_stream_strict(reads: Sequence[bytes], k: int) -> np.ndarray:
    """What kcgpu_add_read builds in a staging block (strict base table everywhere)."""
    return pack_stream(reads, k, simd_rule=False)


def make_genome_reads(rng: np.random.Generator, genome_len: int, n_reads: int, *, mean_len: int = 150, jitter: int = 0,
                      sub_rate: float = 0.01, n_rate: float = 0.003, junk_rate: float = 0.0, lower_rate: float = 0.0,
                      repeat: int = 0) -> List[bytes]:
    """Reads drawn from both strands of one random genome (so that k-mers repeat with the depth),
    with substitutions, N, junk bytes and lower case; `repeat` low-complexity reads are appended
    (poly-A, dinucleotide runs: the same k-mer hundreds of times, which exercises the saturating count)."""
    g = ACGT[rng.integers(0, 4, genome_len)]
    reads = []
    for _ in range(n_reads):
        ln = max(mean_len + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0), 0)
        ln = min(ln, genome_len)
        at = int(rng.integers(0, genome_len - ln + 1))
        s = g[at:at + ln].copy()
        if sub_rate:
            m = rng.random(ln) < sub_rate
            s[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
        if rng.random() < 0.5:
            s = COMP[s][::-1].copy()
        if n_rate:
            s[rng.random(ln) < n_rate] = ord("N")
        if junk_rate:
            m = rng.random(ln) < junk_rate
            s[m] = JUNK[rng.integers(0, len(JUNK), int(m.sum()))]
        if lower_rate:
            m = rng.random(ln) < lower_rate
            s[m] = s[m] | 0x20
        reads.append(s.tobytes())
    for i in range(repeat):
        unit = [b"A", b"AC", b"T", b"ACG", b"GT"][i % 5]
        reads.append((unit * 400)[: 300 + 7 * i])
    return reads


# ---------------------------------------------------------------------------------------------
# snp-pattern-gen cases


def make_spg_case(seed: int, k: int, n_snps: int = 400):
    """(fasta bytes, bed bytes) that exercise every branch of snp-pattern-gen: several contigs
    (one shorter than k), multi-line records with a comment on the header, lower case and U,
    N runs, a stretch copied to another contig forward and one reverse-complemented (reference
    k-mers that occur twice), copies that carry the ALT allele (alternative k-mers that occur),
    SNPs at contig ends, on unknown contigs, with alt '-' / 'N' / lower case, duplicated rows."""
    rng = np.random.default_rng(seed)
    lens = {"chrA": 90000, "chrB": 60000, "chrC": 5000, "tiny": max(k - 3, 1)}
    g = {n: ACGT[rng.integers(0, 4, ln)].copy() for n, ln in lens.items()}
    g["chrB"][1000:4000] = g["chrA"][20000:23000]                      # forward duplicate
    g["chrC"][500:2500] = COMP[g["chrA"][40000:42000]][::-1]           # reverse-complement duplicate
    for n in ("chrA", "chrB"):
        for _ in range(6):
            at = int(rng.integers(0, lens[n] - 50))
            g[n][at:at + int(rng.integers(1, 40))] = ord("N")
        at = int(rng.integers(0, lens[n] - 3000))
        g[n][at:at + 2500] |= 0x20                                       # soft-masked stretch
    rows = []
    flank = k // 2
    for i in range(n_snps):
        n = ["chrA", "chrB", "chrC"][int(rng.integers(0, 3))]
        r = rng.random()
        if r < 0.04:
            pos = int(rng.integers(0, flank + 1))                         # too close to the start
        elif r < 0.08:
            pos = lens[n] - 1 - int(rng.integers(0, flank + 1))           # too close to the end
        elif r < 0.30 and n == "chrA":
            pos = int(rng.integers(20000 + flank, 23000 - flank))         # inside the duplicated stretch
        elif r < 0.40 and n == "chrA":
            pos = int(rng.integers(40000 + flank, 42000 - flank))         # inside the reverse-complemented one
        else:
            pos = int(rng.integers(flank, lens[n] - flank))
        ref = chr(g[n][pos])
        alt = "ACGT"[int(rng.integers(0, 4))]
        while alt.upper() == ref.upper():
            alt = "ACGT"[int(rng.integers(0, 4))]
        r2 = rng.random()
        if r2 < 0.03:
            alt = "-"
        elif r2 < 0.05:
            alt = "N"
        elif r2 < 0.10:
            alt = alt.lower()
        elif r2 < 0.20 and flank <= pos < lens[n] - flank:
            # plant the alternative k-mer somewhere else: the SNP must be rejected
            km = g[n][pos - flank:pos - flank + k].copy()
            km[flank] = ord(alt)
            if rng.random() < 0.5:
                km = COMP[km][::-1]
            at = int(rng.integers(50000, 59000 - k))
            g["chrB"][at:at + k] = km
        chrom = n if rng.random() > 0.03 else "chrUn_missing"
        rows.append((chrom, pos, pos + 1, "rs%d" % i, ref, alt))
        if rng.random() < 0.03:
            rows.append((chrom, pos, pos + 1, "rs%d_dup" % i, ref, alt))
    g["chrA"][70000:70050] = np.frombuffer(b"ACGU" * 12 + b"ug", dtype=np.uint8)  # U counts as T
    fa = []
    for n in ("chrA", "tiny", "chrB", "chrC"):
        fa.append(b">%s some comment here\n" % n.encode())
        s = g[n].tobytes()
        fa.extend(s[j:j + 61] + b"\n" for j in range(0, len(s), 61))
    bed = b"".join(b"%s\t%d\t%d\t%s\t%s\t%s\n" % (c.encode(), s0, e0, rs.encode(), r.encode(), a.encode())
                   for c, s0, e0, rs, r, a in rows)
    return b"".join(fa), bed
