"""ctypes binding of the counting mode of libvafgpu.so (include/kcgpu.h): the kc-c4 path.

The library is the product; this file only marshals arguments.  There is no Python or CPU
implementation of the counting path here: if libvafgpu.so is missing, or no B200 is visible,
construction of a Counter raises.

Reference behaviour behind the calls (paths relative to the reference checkout):
  Counter.add_read / count_device   kc-c4.c:74-90,116-128,133-180
  Counter.histogram                 kc-c4.c:186-215
  format_histogram                  kc-c4.c:232-233
  owner_of                          kc-c4.c:66 (partition by hash suffix), n owners instead of 2^p
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

import vafgpu
from vafgpu import VafGpuError

MAX_OWNERS = 16
IPC_HANDLE_BYTES = 64
NO_LISTS = (1 << 64) - 1  # list_slots: every k-mer straight to the table

EXPORTS = (
    "kcgpu_device_count", "kcgpu_create", "kcgpu_add_read", "kcgpu_producer_create", "kcgpu_producer_add_read",
    "kcgpu_producer_flush", "kcgpu_producer_destroy", "kcgpu_submit_stream", "kcgpu_count_device", "kcgpu_extract_device",
    "kcgpu_insert_device", "kcgpu_table", "kcgpu_ipc_export", "kcgpu_ipc_open", "kcgpu_set_owners",
    "kcgpu_link", "kcgpu_sync", "kcgpu_flush", "kcgpu_histogram", "kcgpu_reset", "kcgpu_destroy", "kcgpu_strerror",
    "kcgpu_hash64", "kcgpu_create_filtered", "kcgpu_set_pass", "kcgpu_histogram1024",
)
PASS_COUNT, PASS_CLAIM, PASS_LOOKUP = 0, 1, 2


class Stats(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_blocks", C.c_uint64), ("n_kmers", C.c_uint64),
        ("n_distinct", C.c_uint64), ("n_overflow", C.c_uint64), ("n_dropped", C.c_uint64), ("n_direct", C.c_uint64),
        ("n_flushes", C.c_uint64), ("table_slots", C.c_uint64), ("list_slots", C.c_uint64), ("flush_bytes", C.c_uint64),
        ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
    ]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


_declared = False


def load_library() -> C.CDLL:
    """libvafgpu.so with the prototypes of include/kcgpu.h declared."""
    global _declared
    lib = vafgpu.load_library()
    if _declared:
        return lib
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    lib.kcgpu_device_count.restype = C.c_int
    lib.kcgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int]
    lib.kcgpu_add_read.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.kcgpu_submit_stream.argtypes = [vp, vp, C.c_size_t]
    lib.kcgpu_producer_create.argtypes = [vp, C.POINTER(vp)]
    lib.kcgpu_producer_add_read.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.kcgpu_producer_flush.argtypes = [vp]
    lib.kcgpu_producer_destroy.argtypes = [vp]
    lib.kcgpu_count_device.argtypes = [vp, vp, C.c_size_t, vp]
    lib.kcgpu_extract_device.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, vp, vp]
    lib.kcgpu_insert_device.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
    lib.kcgpu_table.argtypes = [vp, C.POINTER(vp), u64p]
    lib.kcgpu_ipc_export.argtypes = [vp, vp]
    lib.kcgpu_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.kcgpu_set_owners.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.kcgpu_link.argtypes = [C.POINTER(vp), C.c_int]
    lib.kcgpu_sync.argtypes = [vp]
    lib.kcgpu_flush.argtypes = [vp]
    lib.kcgpu_histogram.argtypes = [vp, u64p, C.POINTER(Stats)]
    lib.kcgpu_reset.argtypes = [vp]
    lib.kcgpu_create_filtered.argtypes = [C.POINTER(vp), C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int]
    lib.kcgpu_set_pass.argtypes = [vp, C.c_int]
    lib.kcgpu_histogram1024.argtypes = [vp, u64p, C.c_int, C.c_int, C.POINTER(Stats)]
    for name in EXPORTS:
        if name not in ("kcgpu_destroy", "kcgpu_strerror", "kcgpu_hash64"):
            getattr(lib, name).restype = C.c_int
    lib.kcgpu_destroy.argtypes = [vp]
    lib.kcgpu_destroy.restype = None
    lib.kcgpu_strerror.argtypes = [vp]
    lib.kcgpu_strerror.restype = C.c_char_p
    lib.kcgpu_hash64.argtypes = [C.c_uint64, C.c_int]
    lib.kcgpu_hash64.restype = C.c_uint64
    _declared = True
    return lib


def hash64(key: int, k: int) -> int:
    return int(load_library().kcgpu_hash64(key, k))


def owner_of(hashed: np.ndarray, n_parts: int) -> np.ndarray:
    """Which of n_parts tables a hashed k-mer belongs to (what the kernels compute)."""
    return (np.asarray(hashed, dtype=np.uint64) % np.uint64(n_parts)).astype(np.int64)


def format_histogram(hist: Sequence[int]) -> str:
    """The 255 lines kc-c4 prints (kc-c4.c:232-233)."""
    return "".join(f"{i}\t{int(hist[i])}\n" for i in range(1, 256))


def format_histogram1024(hist: Sequence[int]) -> str:
    """The 1023 lines yak-count prints (yak-count.c:503)."""
    return "".join(f"{i}\t{int(hist[i])}\n" for i in range(1, 1024))


class Counter:
    """One k-mer table with its region lists on one device."""

    def __init__(self, k: int, table_slots: int = 0, block_bytes: int = 0, device: int = 0, list_slots: int = 0,
                 bloom_bits: int = 0, bloom_hashes: int = 0):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.kcgpu_create_filtered(C.byref(self._ctx), k, table_slots, list_slots, block_bytes, device,
                                             bloom_bits, bloom_hashes)
        if rc:
            raise VafGpuError(rc, self._lib.kcgpu_strerror(None).decode())
        self.k = k
        self.device = device

    def _check(self, rc: int) -> None:
        if rc:
            raise VafGpuError(rc, self._lib.kcgpu_strerror(self._ctx).decode())

    def add_read(self, seq: bytes) -> None:
        self._check(self._lib.kcgpu_add_read(self._ctx, seq, len(seq)))

    def submit_stream(self, host_ptr: int, n_bytes: int) -> None:
        """reads separated by '\\n' in host memory (page-locked: copied without staging)"""
        self._check(self._lib.kcgpu_submit_stream(self._ctx, host_ptr, n_bytes))

    def count_device(self, d_ptr: int, n_bytes: int, stream: Optional[int] = None) -> None:
        self._check(self._lib.kcgpu_count_device(self._ctx, d_ptr, n_bytes, stream))

    def extract_device(self, d_ptr: int, n_bytes: int, n_parts: int, d_keys: int, cap_per_part: int,
                       d_part_counts: int, stream: Optional[int] = None) -> None:
        self._check(self._lib.kcgpu_extract_device(self._ctx, d_ptr, n_bytes, n_parts, d_keys, cap_per_part,
                                                   d_part_counts, stream))

    def insert_device(self, d_keys: int, n: int, n_parts: int, stream: Optional[int] = None) -> None:
        self._check(self._lib.kcgpu_insert_device(self._ctx, d_keys, n, n_parts, stream))

    def table(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.kcgpu_table(self._ctx, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.kcgpu_ipc_export(self._ctx, buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._check(self._lib.kcgpu_ipc_open(self._ctx, handle, C.byref(p)))
        return int(p.value)

    def set_owners(self, my_part: int, tables: Sequence[Optional[int]]) -> None:
        arr = (C.c_void_p * len(tables))(*[C.c_void_p(t) if t else C.c_void_p() for t in tables])
        self._check(self._lib.kcgpu_set_owners(self._ctx, len(tables), my_part, arr))

    def sync(self) -> None:
        self._check(self._lib.kcgpu_sync(self._ctx))

    def flush(self) -> None:
        self._check(self._lib.kcgpu_flush(self._ctx))

    def histogram(self) -> Tuple[np.ndarray, dict]:
        hist = np.zeros(256, dtype=np.uint64)
        st = Stats()
        self._check(self._lib.kcgpu_histogram(self._ctx, hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(st)))
        return hist, st.as_dict()

    def set_pass(self, which: int) -> None:
        """PASS_COUNT / PASS_CLAIM / PASS_LOOKUP: what the insert step does from now on (yak-count's passes)"""
        self._check(self._lib.kcgpu_set_pass(self._ctx, which))

    def histogram1024(self, min_count: int = 1, max_count: int = 1023) -> Tuple[np.ndarray, dict]:
        hist = np.zeros(1024, dtype=np.uint64)
        st = Stats()
        self._check(self._lib.kcgpu_histogram1024(self._ctx, hist.ctypes.data_as(C.POINTER(C.c_uint64)), min_count, max_count,
                                                  C.byref(st)))
        return hist, st.as_dict()

    def reset(self) -> None:
        self._check(self._lib.kcgpu_reset(self._ctx))

    def close(self) -> None:
        if self._ctx:
            self._lib.kcgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def link(counters: Sequence[Counter]) -> None:
    """Contexts of one process, one per device: every kernel adds to the owner's table over NVLink."""
    lib = load_library()
    arr = (C.c_void_p * len(counters))(*[c._ctx for c in counters])
    rc = lib.kcgpu_link(arr, len(counters))
    if rc:
        raise VafGpuError(rc, lib.kcgpu_strerror(counters[0]._ctx).decode())
