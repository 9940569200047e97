"""ctypes binding of librcx.so (include/rcx.h) — plumbing only.

The product is the C-ABI library and the `rcorrector` host binary; this module
lets tests, bench.py and torch.distributed launchers call the same entry points
with numpy arrays or raw device pointers.  It never computes anything itself:
if librcx.so is missing or no GPU is visible, calls raise (no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCX_LIB", os.path.join(_HERE, "librcx.so"))  # RCX_LIB: profiling build (make prof)

SINGLE, INTERLEAVED = 0, 1

# every symbol include/rcx.h declares
SYMBOLS = [
    "rcx_abi_version", "rcx_last_error", "rcx_create", "rcx_destroy", "rcx_table_build", "rcx_table_build_multi", "rcx_table_lookup",
    "rcx_table_info", "rcx_count_kmers_begin", "rcx_count_kmers_add", "rcx_count_kmers_finish", "rcx_count_kmers_export", "rcx_set_params", "rcx_correct_batch", "rcx_correct_batch_compact", "rcx_correct_batch_device", "rcx_sync",
    "rcx_scan_batch", "rcx_get_stats", "rcx_debug_counters", "rcx_measure_random_read", "rcx_stream", "rcx_host_alloc", "rcx_host_free", "rcx_estimate_error_rate",
]


class RcxParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("max_fix_per_k", C.c_int32), ("error_rate", C.c_double), ("bad_qual", C.c_int32)]


class RcxStats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("bases", C.c_uint64), ("scan_probes", C.c_uint64),
                ("search_probes", C.c_uint64), ("rescan_probes", C.c_uint64), ("search_nodes", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("ms_scan", C.c_float), ("ms_correct", C.c_float),
                ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("ms_triage", C.c_float), ("ms_prep", C.c_float), ("ms_dfs", C.c_float), ("ms_finalize", C.c_float),
                ("work_items", C.c_uint64), ("second_pass", C.c_uint64), ("heavy_items", C.c_uint64), ("vheavy_items", C.c_uint64), ("rescan_windows", C.c_uint64), ("unsupported", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class RcxError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen librcx.so; raises if it has not been built (python __graft_entry__.py / make -C rcorrector_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RcxError("librcx.so not built: run `make -C rcorrector_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32
    lib.rcx_abi_version.restype = C.c_int
    lib.rcx_last_error.restype = C.c_char_p
    lib.rcx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.rcx_destroy.argtypes = [vp]
    lib.rcx_destroy.restype = None
    lib.rcx_table_build.argtypes = [vp, C.c_int, vp, vp, u64, C.c_int]
    lib.rcx_table_build_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp, vp, u64]
    lib.rcx_table_lookup.argtypes = [vp, vp, vp, u64]
    lib.rcx_table_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    lib.rcx_count_kmers_begin.argtypes = [vp, C.c_int, u64]
    lib.rcx_count_kmers_add.argtypes = [vp, i64, vp, vp]
    lib.rcx_count_kmers_finish.argtypes = [vp, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    lib.rcx_count_kmers_export.argtypes = [vp, vp, vp, u64]
    lib.rcx_set_params.argtypes = [vp, C.POINTER(RcxParams)]
    lib.rcx_correct_batch.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rcx_correct_batch_compact.argtypes = [vp, C.c_int, i64, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
    lib.rcx_correct_batch_device.argtypes = [vp, C.c_int, i64, vp, vp, vp, u64, i32, vp, vp, vp, vp]
    lib.rcx_sync.argtypes = [vp]
    lib.rcx_scan_batch.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.rcx_get_stats.argtypes = [vp, C.POINTER(RcxStats)]
    lib.rcx_debug_counters.argtypes = [vp, vp, C.c_int]
    lib.rcx_measure_random_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.rcx_stream.argtypes = [vp]
    lib.rcx_stream.restype = vp
    lib.rcx_host_alloc.argtypes = [u64]
    lib.rcx_host_alloc.restype = vp
    lib.rcx_host_free.argtypes = [vp]
    lib.rcx_host_free.restype = None
    lib.rcx_estimate_error_rate.argtypes = [vp, vp, u64, C.c_double, C.POINTER(C.c_double)]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host)
        return a.data_ptr()
    raise TypeError(type(a))


class Corrector:
    """One GPU context: the table in HBM plus the per-batch hot path (mirrors what main.cpp sets up
    before its correction loop: Store kmers, kmerLength, MAX_FIX_PER_K, ERROR_RATE, badQualityThreshold)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        self._check(self.lib.rcx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.k = None

    def _check(self, rc):
        if rc != 0:
            raise RcxError("rcx error %d: %s" % (rc, self.lib.rcx_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rcx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- table
    def build_table(self, k, codes, counts, n=None, on_device=False):
        if not on_device:
            codes = np.ascontiguousarray(codes, dtype=np.uint64)
            counts = np.ascontiguousarray(counts, dtype=np.int32)
            n = len(codes)
        self._check(self.lib.rcx_table_build(self.h, k, _ptr(codes), _ptr(counts), n, 1 if on_device else 0))
        self.k = k

    def count_kmers(self, k, batches, min_count=2, capacity_hint=0):
        """count the canonical k-mers of `batches` (iterable of (seq, off) host arrays) on the GPU, keep those seen
        at least min_count times and build the count table from them.  Returns (codes, counts, distinct, instances)."""
        self._check(self.lib.rcx_count_kmers_begin(self.h, k, capacity_hint))
        for seq, off in batches:
            seq = np.ascontiguousarray(seq, dtype=np.uint8)
            off = np.ascontiguousarray(off, dtype=np.uint32)
            self._check(self.lib.rcx_count_kmers_add(self.h, len(off) - 1, _ptr(seq), _ptr(off)))
        kept, distinct, inst = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.rcx_count_kmers_finish(self.h, min_count, C.byref(kept), C.byref(distinct), C.byref(inst)))
        codes = np.zeros(kept.value, dtype=np.uint64)
        counts = np.zeros(kept.value, dtype=np.int32)
        self._check(self.lib.rcx_count_kmers_export(self.h, _ptr(codes), _ptr(counts), kept.value))
        self.k = k
        return codes, counts, distinct.value, inst.value

    def table_info(self):
        e, s, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.rcx_table_info(self.h, C.byref(e), C.byref(s), C.byref(b)))
        return {"entries": e.value, "slots": s.value, "bytes": b.value}

    def lookup(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(len(codes), dtype=np.int32)
        self._check(self.lib.rcx_table_lookup(self.h, _ptr(codes), _ptr(out), len(codes)))
        return out

    def estimate_error_rate(self, codes, wk=0.95):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        out = C.c_double()
        self._check(self.lib.rcx_estimate_error_rate(self.h, _ptr(codes), len(codes), wk, C.byref(out)))
        return out.value

    def set_params(self, k, error_rate, bad_qual, max_fix_per_k=4):
        p = RcxParams(k, max_fix_per_k, error_rate, bad_qual)
        self._check(self.lib.rcx_set_params(self.h, C.byref(p)))

    # ---- hot path
    def correct(self, mode, seq, qual, off, out=None):
        """Host arrays in, host arrays out (H2D + kernels + D2H inside the call)."""
        n = len(off) - 1
        if out is None:
            out = (np.empty(len(seq), dtype=np.uint8), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32),
                   np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32))
        so, corr, l, m, h = out
        self._check(self.lib.rcx_correct_batch(self.h, mode, n, _ptr(seq), _ptr(qual), _ptr(off), _ptr(so), _ptr(corr),
                                               _ptr(l), _ptr(m), _ptr(h)))
        return so, corr, l, m, h

    def correct_compact(self, mode, seq, qual_bits, off, out=None, has_qual=True):
        """Compact host interface: one quality bit per base in (see quality_bits()), applied edits out instead of sequences.
        Returns (correction, l, m, h, edits[:n_edits])."""
        n = len(off) - 1
        if out is None:
            out = (np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32),
                   np.zeros(max(len(seq) // 4, 1024), dtype=np.uint16))
        corr, l, m, h, edits = out
        ne = C.c_uint64()
        self._check(self.lib.rcx_correct_batch_compact(self.h, mode, n, _ptr(seq), _ptr(qual_bits), 1 if has_qual else 0, _ptr(off),
                                                       _ptr(corr), _ptr(l), _ptr(m), _ptr(h), _ptr(edits), len(edits), C.byref(ne)))
        return corr, l, m, h, edits[: ne.value]

    def correct_device(self, mode, n, seq, qual, off, total_bases, max_len, corr, l, m, h):
        """Device pointers / torch CUDA tensors; seq corrected in place; asynchronous (call sync())."""
        self._check(self.lib.rcx_correct_batch_device(self.h, mode, n, _ptr(seq), _ptr(qual), _ptr(off), total_bases,
                                                      max_len, _ptr(corr), _ptr(l), _ptr(m), _ptr(h)))

    def sync(self):
        self._check(self.lib.rcx_sync(self.h))

    def scan(self, seq, off):
        n = len(off) - 1
        counts = np.zeros(len(seq), dtype=np.int32)
        thr = np.zeros(n, dtype=np.int32)
        self._check(self.lib.rcx_scan_batch(self.h, n, _ptr(seq), _ptr(off), _ptr(counts), _ptr(thr)))
        return counts, thr

    def stats(self):
        s = RcxStats()
        self._check(self.lib.rcx_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def measure_random_read(self, sector_bytes=32):
        out = C.c_double()
        self._check(self.lib.rcx_measure_random_read(self.h, sector_bytes, C.byref(out)))
        return out.value

    def debug_counters(self, n=16):
        out = np.zeros(n, dtype=np.uint64)
        self._check(self.lib.rcx_debug_counters(self.h, _ptr(out), n))
        return out

    def stream(self):
        return self.lib.rcx_stream(self.h)


def quality_bits(qual, bad_qual):
    """The quality bit plane rcx_correct_batch_compact takes: bit (i & 7) of byte i >> 3 = (signed char)qual[i] > bad_qual."""
    q = np.ascontiguousarray(qual, dtype=np.uint8).view(np.int8)
    return np.packbits(q > bad_qual, bitorder="little")


def apply_edits(seq, off, corr, edits):
    """What a caller of the compact interface does with its own copy of the reads: record r owns corr[r] entries of
    `edits` (when positive) after those of the records before it; entry = position | base << 10."""
    out = np.array(seq, dtype=np.uint8, copy=True)
    c = np.maximum(np.asarray(corr, dtype=np.int64), 0)
    start = np.repeat(np.asarray(off[:-1], dtype=np.int64), c)
    e = np.asarray(edits, dtype=np.uint16).astype(np.int64)
    assert len(e) == len(start)
    out[start + (e & 1023)] = np.frombuffer(b"ACGT", dtype=np.uint8)[e >> 10]
    return out
