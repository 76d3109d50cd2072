"""ctypes binding of libvafgpu.so (include/vafgpu.h) plus the small host-side mirror of the
reference's pattern handling that tests and bench.py need.

The library is the product; this file only marshals arguments.  There is no Python or CPU
implementation of the counting path here: if libvafgpu.so is missing, or no B200 is
visible, construction of an Engine raises.

Reference behaviour mirrored by the helpers (paths relative to the reference checkout):
  load_patterns      vaf-counter.c:149-184
  build_key_list     vaf-counter.c:198-252 (first insert wins, collisions counted)
  write_vaf          vaf-counter.c:654-680
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAFGPU_LIB") or os.path.join(_HERE, "libvafgpu.so")  # VAFGPU_LIB: a differently tuned build (development)

F_REFERENCE_RECIPE = 1
F_HOST_MERGE = 2
F_STRICT_BYTES = 4

OK, EINVAL, ENOGPU, ECUDA, ENOMEM, ENCCL, ESTATE = 0, -1, -2, -3, -4, -5, -6

EXPORTS = (
    "vafgpu_create", "vafgpu_add_read", "vafgpu_submit_stream", "vafgpu_count_device",
    "vafgpu_finish", "vafgpu_reset", "vafgpu_destroy", "vafgpu_strerror", "vafgpu_plan",
    "vafgpu_canonicalise_read", "vafgpu_version",
    "vafgpu_producer_create", "vafgpu_producer_add_read", "vafgpu_producer_flush",
    "vafgpu_producer_destroy", "vafgpu_export_counters", "vafgpu_attach_counters",
)


class Stats(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_blocks", C.c_uint64),
        ("n_bytes", C.c_uint64), ("n_candidates", C.c_uint64), ("n_hits", C.c_uint64),
        ("n_kmers", C.c_uint64), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
        ("n_devices", C.c_int), ("anchor_stride", C.c_int), ("anchor_len", C.c_int),
        ("filter_bytes", C.c_uint32), ("table_slots", C.c_uint32),
        ("filter_canon", C.c_int), ("lookup_deferred", C.c_int), ("kernel_threads", C.c_int),
        ("filter2_bytes", C.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


class VafGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vafgpu error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libvafgpu.so and declare the prototypes of include/vafgpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `make -C {_HERE}` (or __graft_entry__.build()); "
            "there is no fallback implementation")
    lib = C.CDLL(path)
    u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    lib.vafgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, u64p, u32p, C.c_uint32, C.c_uint32,
                                  C.c_size_t, C.c_int, C.c_int, C.c_uint]
    lib.vafgpu_create.restype = C.c_int
    lib.vafgpu_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.vafgpu_add_read.restype = C.c_int
    lib.vafgpu_submit_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
    lib.vafgpu_submit_stream.restype = C.c_int
    lib.vafgpu_count_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.vafgpu_count_device.restype = C.c_int
    lib.vafgpu_finish.argtypes = [C.c_void_p, u32p, C.POINTER(Stats)]
    lib.vafgpu_finish.restype = C.c_int
    lib.vafgpu_reset.argtypes = [C.c_void_p]
    lib.vafgpu_reset.restype = C.c_int
    lib.vafgpu_destroy.argtypes = [C.c_void_p]
    lib.vafgpu_destroy.restype = None
    lib.vafgpu_strerror.argtypes = [C.c_void_p]
    lib.vafgpu_strerror.restype = C.c_char_p
    lib.vafgpu_plan.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vafgpu_plan.restype = C.c_int
    lib.vafgpu_canonicalise_read.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]
    lib.vafgpu_canonicalise_read.restype = None
    lib.vafgpu_version.restype = C.c_char_p
    lib.vafgpu_producer_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.vafgpu_producer_create.restype = C.c_int
    lib.vafgpu_producer_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.vafgpu_producer_add_read.restype = C.c_int
    lib.vafgpu_producer_flush.argtypes = [C.c_void_p]
    lib.vafgpu_producer_flush.restype = C.c_int
    lib.vafgpu_producer_destroy.argtypes = [C.c_void_p]
    lib.vafgpu_producer_destroy.restype = C.c_int
    lib.vafgpu_export_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.vafgpu_export_counters.restype = C.c_int
    lib.vafgpu_attach_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.vafgpu_attach_counters.restype = C.c_int
    _lib = lib
    return lib


def plan(k: int) -> Tuple[int, int]:
    s, l = C.c_int(), C.c_int()
    rc = load_library().vafgpu_plan(k, C.byref(s), C.byref(l))
    if rc:
        raise VafGpuError(rc, f"k={k} is outside 1..31")
    return s.value, l.value


def canonicalise_read(seq: bytes, simd_rule: bool = True) -> bytes:
    out = C.create_string_buffer(len(seq))
    load_library().vafgpu_canonicalise_read(seq, len(seq), out, 1 if simd_rule else 0)
    return out.raw


class Engine:
    """One vafgpu context: tables replicated on n_devices GPUs, counters accumulated there."""

    def __init__(self, k: int, keys: np.ndarray, vals: np.ndarray, n_patterns: int, *,
                 block_bytes: int = 0, n_buffers: int = 0, n_devices: int = 0, flags: int = 0):
        self._lib = load_library()
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        if keys.shape != vals.shape:
            raise ValueError("keys and vals differ in length")
        self._h = C.c_void_p()
        self.k, self.n_patterns = k, n_patterns
        rc = self._lib.vafgpu_create(
            C.byref(self._h), k, keys.ctypes.data_as(C.POINTER(C.c_uint64)),
            vals.ctypes.data_as(C.POINTER(C.c_uint32)), keys.size, n_patterns, block_bytes,
            n_buffers, n_devices, flags)
        if rc:
            self._h = C.c_void_p()
            raise VafGpuError(rc, self._lib.vafgpu_strerror(None).decode())

    def _check(self, rc: int) -> None:
        if rc:
            raise VafGpuError(rc, self._lib.vafgpu_strerror(self._h).decode())

    def add_read(self, seq: bytes) -> None:
        self._check(self._lib.vafgpu_add_read(self._h, seq, len(seq)))

    def producer(self) -> "Producer":
        """A producer for one reader thread (vafgpu_producer_*); flush or close it before finish()."""
        return Producer(self)

    def submit_stream(self, buf, n_reads: int = 0, n_bases: int = 0) -> None:
        """buf: bytes, a numpy uint8 array, or (address, n_bytes) of host memory in stream form."""
        if isinstance(buf, tuple):
            ptr, n = buf
        elif isinstance(buf, np.ndarray):
            ptr, n = buf.ctypes.data, buf.nbytes
        else:
            self._keep = buf
            ptr, n = C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf)
        self._check(self._lib.vafgpu_submit_stream(self._h, ptr, n, n_reads, n_bases))

    def count_device(self, d_ptr: int, n_bytes: int, *, device: int = 0, d_counts: int = 0,
                     stream: int = 0) -> None:
        self._check(self._lib.vafgpu_count_device(self._h, device, d_ptr, n_bytes,
                                                  d_counts or None, stream or None))

    def finish(self) -> Tuple[np.ndarray, dict]:
        counts = np.zeros(2 * max(self.n_patterns, 1), dtype=np.uint32)
        st = Stats()
        self._check(self._lib.vafgpu_finish(self._h, counts.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(st)))
        return counts[: 2 * self.n_patterns], st.as_dict()

    def reset(self) -> None:
        self._check(self._lib.vafgpu_reset(self._h))

    def export_counters(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this engine's counter vector, for attach_counters elsewhere."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.vafgpu_export_counters(self._h, buf, 64))
        return buf.raw

    def attach_counters(self, handle: bytes) -> None:
        """From now on this engine's kernels add into the exporting engine's vector (another process)."""
        buf = C.create_string_buffer(bytes(handle), 64)
        self._check(self._lib.vafgpu_attach_counters(self._h, buf, 64))

    def close(self) -> None:
        if self._h:
            self._lib.vafgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Producer:
    """One reader thread's handle on an Engine (parallel ingest)."""

    def __init__(self, engine: Engine):
        self._e = engine
        self._p = C.c_void_p()
        engine._check(engine._lib.vafgpu_producer_create(engine._h, C.byref(self._p)))

    def add_read(self, seq: bytes) -> None:
        self._e._check(self._e._lib.vafgpu_producer_add_read(self._p, seq, len(seq)))

    def flush(self) -> None:
        self._e._check(self._e._lib.vafgpu_producer_flush(self._p))

    def close(self) -> None:
        if self._p:
            p, self._p = self._p, C.c_void_p()
            self._e._check(self._e._lib.vafgpu_producer_destroy(p))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---------------------------------------------------------------------------------------------
# host-side mirror of the reference's pattern handling


@dataclass
class Pattern:
    chr: str
    start: int
    end: int
    rsid: str
    ref: str
    alt: str
    ref_kmer: str
    alt_kmer: str


def load_patterns(path: str) -> List[Pattern]:
    """Eight white-space separated fields per record; the first malformed record ends the load
    (fscanf loop of vaf-counter.c:164-180; field width caps 255/255/127 apply)."""
    with open(path, "rb") as fh:
        tok = fh.read().split()
    out: List[Pattern] = []
    i = 0
    while i + 8 <= len(tok):
        t = tok[i:i + 8]
        try:
            start, end = int(t[1]), int(t[2])
        except ValueError:
            break
        if len(t[0]) > 255 or len(t[3]) > 255 or len(t[4]) != 1 or len(t[5]) != 1 \
                or len(t[6]) > 127 or len(t[7]) > 127:
            break  # fscanf would split the token differently; the reference's files never do this
        out.append(Pattern(t[0].decode(), start, end, t[3].decode(), t[4].decode(), t[5].decode(),
                           t[6].decode(), t[7].decode()))
        i += 8
    return out


_CODE = np.full(256, 4, dtype=np.uint8)
for _b, _c in ((0, 0), (1, 1), (2, 2), (3, 3)):
    _CODE[_b] = _c
for _ch, _c in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _CODE[ord(_ch)] = _c
    _CODE[ord(_ch.lower())] = _c


def canonical_kmer(s: str, k: int) -> Optional[int]:
    """Canonical k-mer of the first k characters in the reference encoding, None if any of
    them is not a base (vaf-counter.c:117-146)."""
    b = s.encode()[:k]
    if len(b) < k:
        # the reference reads the NUL terminator, which is not a base... except that byte 0
        # IS code 0 in its table; patterns shorter than k do not occur in practice
        b = b + b"\0" * (k - len(b))
    f = r = 0
    for ch in b:
        c = int(_CODE[ch])
        if c > 3:
            return None
        f = (f << 2) | c
        r = (r >> 2) | ((3 - c) << (2 * (k - 1)))
    return min(f, r)


def build_key_list(patterns: Sequence[Pattern], k: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(keys, vals, n_collisions): canonical(ref) -> i<<1, canonical(alt) -> i<<1|1 in file
    order; a key already present keeps its first value (vaf-counter.c:218-244)."""
    seen = {}
    n_coll = 0
    for i, p in enumerate(patterns):
        for alt, s in ((0, p.ref_kmer), (1, p.alt_kmer)):
            c = canonical_kmer(s, k)
            if c is None:
                continue
            if c in seen:
                n_coll += 1
            else:
                seen[c] = (i << 1) | alt
    keys = np.fromiter(seen.keys(), dtype=np.uint64, count=len(seen))
    vals = np.fromiter(seen.values(), dtype=np.uint32, count=len(seen))
    return keys, vals, n_coll


def format_vaf(patterns: Sequence[Pattern], counts: Iterable[int]) -> str:
    """The .vaf text of vaf-counter.c:654-680, byte for byte."""
    c = np.asarray(list(counts) if not isinstance(counts, np.ndarray) else counts, dtype=np.uint64)
    n = len(patterns)
    tot = int(c[: 2 * n].sum())
    lines = ["# Average depth: %.2f\n" % (tot / (n if n > 0 else 1)),
             "CHR\tPOS\tRSID\tREF\tALT\tREF_COUNT\tALT_COUNT\tTOTAL_COUNT\tVAF\n"]
    for i, p in enumerate(patterns):
        r, a = int(c[2 * i]), int(c[2 * i + 1])
        t = (r + a) & 0xFFFFFFFF
        vaf = (a / t) if t > 0 else 0.0
        lines.append("%s\t%d\t%s\t%s\t%s\t%u\t%u\t%u\t%.4f\n" % (p.chr, p.start, p.rsid, p.ref, p.alt, r, a, t, vaf))
    return "".join(lines)


def write_vaf(path: str, patterns: Sequence[Pattern], counts) -> None:
    with open(path, "w") as fh:
        fh.write(format_vaf(patterns, counts))
