"""Shared test helpers: ctypes binding of the CPU oracle, FASTQ <-> batch
conversion that follows the reference's reader/writer byte for byte
(Reads.h:224-266 NextWithBuffer, Reads.h:360-421 OutputBatch), and runners
for the tools and the reference binary.

TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke() import this module.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "librcx_oracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
SYNTH = os.path.join(ROOT, "tools", "bin", "rcx_synth")
KMERCOUNT = os.path.join(ROOT, "tools", "bin", "rcx_kmercount")
GOLDEN = os.path.join(ROOT, "tests", "golden")

SINGLE, INTERLEAVED = 0, 1


class OrcParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("max_fix_per_k", C.c_int32), ("error_rate", C.c_double), ("bad_qual", C.c_int32)]


_orc = None


def oracle():
    """Load oracle/librcx_oracle.so (built by oracle/Makefile / __graft_entry__.build())."""
    global _orc
    if _orc is not None:
        return _orc
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "librcx_oracle.so"])
    lib = C.CDLL(ORACLE_SO)
    vp, i32p, u32p, u64p, cp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_char_p
    lib.orc_table_new.restype = vp
    lib.orc_table_new.argtypes = [C.c_int]
    lib.orc_table_free.argtypes = [vp]
    lib.orc_table_size.restype = C.c_uint64
    lib.orc_table_size.argtypes = [vp]
    lib.orc_table_put.argtypes = [vp, vp, vp, C.c_uint64]
    lib.orc_table_load_dump.restype = C.c_int64
    lib.orc_table_load_dump.argtypes = [vp, cp]
    lib.orc_table_count.restype = C.c_int32
    lib.orc_table_count.argtypes = [vp, C.c_uint64]
    lib.orc_estimate_error_rate.restype = C.c_double
    lib.orc_estimate_error_rate.argtypes = [vp, cp, C.c_double]
    lib.orc_bad_quality.restype = C.c_int32
    lib.orc_bad_quality.argtypes = [vp, vp, C.c_int64]
    lib.orc_scan_counts.restype = C.c_int32
    lib.orc_scan_counts.argtypes = [vp, vp, C.c_int32, vp]
    lib.orc_strong_threshold.restype = C.c_int32
    lib.orc_strong_threshold.argtypes = [vp, vp, C.c_int32]
    lib.orc_correct_batch.argtypes = [vp, C.POINTER(OrcParams), C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    _orc = lib
    return lib


class OracleTable:
    def __init__(self, k):
        self.k = k
        self.lib = oracle()
        self.h = self.lib.orc_table_new(k)

    def put(self, codes, counts):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        self.lib.orc_table_put(self.h, codes.ctypes.data, counts.ctypes.data, len(codes))

    def load_dump(self, path):
        n = self.lib.orc_table_load_dump(self.h, path.encode())
        if n < 0:
            raise IOError(path)
        return n

    def error_rate(self, dump_path, wk=0.95):
        return self.lib.orc_estimate_error_rate(self.h, dump_path.encode(), wk)

    def __len__(self):
        return self.lib.orc_table_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_table_free(self.h)
            self.h = None


class Batch:
    """Reads concatenated without separators; record r = [off[r], off[r+1])."""

    def __init__(self, ids, seq, qual, off):
        self.ids = ids
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.qual = np.ascontiguousarray(qual, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.uint32)

    @property
    def n(self):
        return len(self.off) - 1

    def slice(self, lo, hi):
        o = self.off[lo : hi + 1].astype(np.int64)
        return Batch(self.ids[lo:hi] if self.ids else None, self.seq[o[0] : o[-1]], self.qual[o[0] : o[-1]], (o - o[0]).astype(np.uint32))


def read_fastq(path):
    """Reads.h:224-266: 4 lines per record; the newline of the quality line is
    stripped at the SEQUENCE line's length; text after '+' is dropped."""
    ids, seqs, quals = [], [], []
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\n")
            f.readline()
            q = f.readline()
            q = q[: len(s)] if len(q) > len(s) and q[len(s) : len(s) + 1] == b"\n" else q.rstrip(b"\n")
            ids.append(h.rstrip(b"\n"))
            seqs.append(s)
            quals.append(q)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum(lens)
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    qual = np.frombuffer(b"".join(quals), dtype=np.uint8)
    assert len(qual) == len(seq), "quality and sequence lengths differ"
    return Batch(ids, seq, qual, off)


def interleave(b1, b2):
    assert b1.n == b2.n
    n = b1.n
    l1 = np.diff(b1.off.astype(np.int64))
    l2 = np.diff(b2.off.astype(np.int64))
    lens = np.empty(2 * n, dtype=np.int64)
    lens[0::2] = l1
    lens[1::2] = l2
    off = np.zeros(2 * n + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    seq = np.empty(off[-1], dtype=np.uint8)
    qual = np.empty(off[-1], dtype=np.uint8)
    ids = []
    for i in range(n):
        for m, b in ((0, b1), (1, b2)):
            r = 2 * i + m
            seq[off[r] : off[r + 1]] = b.seq[b.off[i] : b.off[i + 1]]
            qual[off[r] : off[r + 1]] = b.qual[b.off[i] : b.off[i + 1]]
            ids.append(b.ids[i])
    return Batch(ids, seq, qual, off.astype(np.uint32))


def format_fastq(batch, seq_out, corr, l, m, h, rows=None):
    """Reads.h:360-421 OutputBatch (FASTQ)."""
    out = []
    seq_out = np.asarray(seq_out, dtype=np.uint8)
    rows = range(batch.n) if rows is None else rows
    for r in rows:
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        tag = b"" if corr[r] == 0 else (b" unfixable_error" if corr[r] == -1 else b" cor")
        out.append(batch.ids[r] + b" l:%d m:%d h:%d" % (l[r], m[r], h[r]) + tag + b"\n")
        out.append(seq_out[a:b].tobytes() + b"\n+\n" + batch.qual[a:b].tobytes() + b"\n")
    return b"".join(out)


def oracle_correct(table, batch, k, mode, error_rate, bad_qual, max_fix_per_k=4, threads=8):
    lib = oracle()
    n = batch.n
    p = OrcParams(k, max_fix_per_k, error_rate, bad_qual)
    seq_out = np.empty_like(batch.seq)
    corr = np.zeros(n, dtype=np.int32)
    l = np.zeros(n, dtype=np.int32)
    m = np.zeros(n, dtype=np.int32)
    h = np.zeros(n, dtype=np.int32)
    lib.orc_correct_batch(table.h, C.byref(p), mode, n, batch.seq.ctypes.data, batch.qual.ctypes.data,
                          batch.off.ctypes.data, seq_out.ctypes.data, corr.ctypes.data, l.ctypes.data,
                          m.ctypes.data, h.ctypes.data, threads)
    return seq_out, corr, l, m, h


def oracle_bad_quality(batch):
    return oracle().orc_bad_quality(batch.qual.ctypes.data, batch.off.ctypes.data, batch.n)


def oracle_scan_counts(table, batch):
    """counts[] per base offset (kcnt entries valid per read), and the pair-threshold per read."""
    lib = oracle()
    counts = np.zeros(len(batch.seq) + 1, dtype=np.int32)
    thr = np.zeros(batch.n, dtype=np.int32)
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        lib.orc_scan_counts(table.h, batch.seq.ctypes.data + a, b - a, counts.ctypes.data + 4 * a)
        thr[r] = lib.orc_strong_threshold(table.h, batch.seq.ctypes.data + a, b - a)
    return counts[:-1], thr


def md5(b):
    return hashlib.md5(b).hexdigest()


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def ensure_tools():
    if not (os.path.exists(SYNTH) and os.path.exists(KMERCOUNT)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools")])


def synth(workdir, name, k=23, paired=True, extra=(), count=True, **kw):
    """Run tools/bin/rcx_synth (+ rcx_kmercount for an exact-count dump). Returns dict of paths."""
    ensure_tools()
    p = os.path.join(workdir, name)
    args = [SYNTH, "-k", str(k), "-fq1", p + "_1.fq", "-bin", p]
    if paired:
        args += ["-paired", "-fq2", p + "_2.fq"]
    for key, v in kw.items():
        args += ["-" + key, str(v)]
    args += list(extra)
    run(args)
    files = [p + "_1.fq"] + ([p + "_2.fq"] if paired else [])
    out = {"fq": files, "bin": p, "dump": p + ".dump", "tab": p + ".tab"}
    if count:
        run([KMERCOUNT, "-k", str(k), "-o", out["dump"], "-bin", out["tab"]] + files)
    return out


def load_table_arrays(prefix):
    return np.fromfile(prefix + ".codes", dtype=np.uint64), np.fromfile(prefix + ".counts", dtype=np.int32)


def run_reference(fq, dump, k, outdir, threads=1, extra=(), mode=None, binary=None):
    """Run the reference binary (oracle/_ref/rcorrector). Returns (stderr text, list of output paths)."""
    os.makedirs(outdir, exist_ok=True)
    if mode is None:
        mode = "-p" if len(fq) == 2 else "-r"
    cmd = [binary or REF_BIN, mode] + list(fq) + ["-c", dump, "-k", str(k), "-od", outdir, "-t", str(threads)] + list(extra)
    r = run(cmd)
    outs = []
    for f in fq:
        base = os.path.basename(f)
        base = base[: base.rfind(".")] if "." in base else base
        outs.append(os.path.join(outdir, base + ".cor.fq"))
    return r.stderr.decode(), outs


def parse_ref_stderr(text):
    """Pull the scalars the reference prints (main.cpp:308,358,364,81-86)."""
    import re

    d = {}
    m = re.search(r"Stored (\d+) kmers", text)
    d["stored"] = int(m.group(1)) if m else None
    m = re.search(r"Weak kmer threshold rate: ([0-9.]+)", text)
    d["rate6"] = m.group(1) if m else None
    m = re.search(r"Bad quality threshold is '(.?)'", text, re.S)
    d["bad_qual"] = ord(m.group(1)) if m and m.group(1) else 0
    m = re.search(r"Processed (\d+) reads\s+Corrected (\d+) bases", text)
    d["reads"], d["bases"] = (int(m.group(1)), int(m.group(2))) if m else (None, None)
    return d


# ----------------------------------------------------------------------------- golden fixtures (tests/golden/make_golden.py)
def golden_cases():
    out = []
    if os.path.isdir(GOLDEN):
        for name in sorted(os.listdir(GOLDEN)):
            if os.path.exists(os.path.join(GOLDEN, name, "meta.json")):
                out.append(name)
    return out


def load_golden(name, workdir):
    """Unpack one fixture into workdir; returns (meta, fastq paths, dump path, expected output bytes list)."""
    import gzip
    import json

    d = os.path.join(GOLDEN, name)
    meta = json.load(open(os.path.join(d, "meta.json")))
    w = os.path.join(workdir, "golden_" + name)
    os.makedirs(w, exist_ok=True)

    def unz(fn):
        dst = os.path.join(w, fn[:-3])
        if not os.path.exists(dst):
            with gzip.open(os.path.join(d, fn), "rb") as g, open(dst, "wb") as f:
                f.write(g.read())
        return dst

    fq = [unz(f) for f in meta["inputs"]]
    dump = unz("reads.dump.gz")
    import gzip as _g

    expected = [_g.open(os.path.join(d, f), "rb").read() for f in meta["outputs"]]
    return meta, fq, dump, expected


def dump_to_arrays(path, k):
    """Parse a Jellyfish-style text dump like main.cpp:294-308 into (forward codes, counts) for rcx_table_build."""
    codes, counts = [], []
    mask = (1 << (2 * k)) - 1
    val = {"A": 0, "C": 1, "G": 2, "T": 3}
    with open(path) as f:
        toks = f.read().split()
    for i in range(0, len(toks) - 1, 2):
        c = int(toks[i][1:])
        code = 0
        ok = True
        for ch in toks[i + 1]:
            v = val.get(ch, -1)
            if v < 0:
                ok = False
                v = 3
            code = ((code << 2) | v) & mask
        codes.append(code)
        counts.append(c if ok else 0)
    return np.array(codes, dtype=np.uint64), np.array(counts, dtype=np.int32)
