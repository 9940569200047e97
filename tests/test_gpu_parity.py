"""GPU parity: the CUDA path through the C ABI (librcx.so) against the CPU oracle
on the same seeded inputs.  Bit-exact: corrected bases, correction code, l/m/h."""
import numpy as np
import pytest

import rcx_testlib as T

pytestmark = pytest.mark.gpu

# name, k, paired, maxcorK, synth kwargs
CASES = [
    ("p100", 23, True, 4, dict(seed=1, ntx=50, reads=20000, len=100, err=0.01)),
    ("s100", 23, False, 4, dict(seed=2, ntx=50, reads=20000, len=100, err=0.01)),
    ("k31deep", 31, True, 8, dict(seed=5, ntx=20, reads=5000, len=150, err=0.04)),
    ("repeat", 23, True, 4, dict(seed=11, ntx=25, reads=30000, len=100, err=0.02, paralog=0.4, repeat=0.4, nrate=0.01)),
    ("maxcor1", 23, True, 1, dict(seed=12, ntx=25, reads=10000, len=100, err=0.02, paralog=0.4, repeat=0.4, nrate=0.01)),
    ("maxcor2", 23, True, 2, dict(seed=13, ntx=25, reads=10000, len=100, err=0.03, paralog=0.5, repeat=0.5, nrate=0.01)),
    ("hicov", 23, True, 4, dict(seed=14, ntx=10, reads=30000, len=100, err=0.01, sigma=1.0)),
    ("k32", 32, True, 4, dict(seed=15, ntx=30, reads=5000, len=120, err=0.02)),
    ("k15", 15, False, 4, dict(seed=16, ntx=30, reads=5000, len=60, err=0.02, txmin=100)),
    ("long250", 25, True, 4, dict(seed=17, ntx=30, reads=4000, len=250, err=0.02, txmin=400)),
]


def _check_compact(gpu, mode, b, bq, exp, has_qual=True):
    """The compact host interface (one quality bit per base in, applied edits out) gives the same reads and tags."""
    from rcorrector_b200.api import apply_edits, quality_bits

    corr, l, m, h, edits = gpu.correct_compact(mode, b.seq, quality_bits(b.qual, bq) if has_qual else None, b.off, has_qual=has_qual)
    assert len(edits) == int(np.maximum(corr, 0).sum())
    got = (apply_edits(b.seq, b.off, corr, edits), corr, l, m, h)
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "compact %s: %d mismatches, first at %d" % (nm, len(bad), bad[0])


def _load(workdir, name, k, paired, kw):
    d = T.synth(workdir, name, k=k, paired=paired, **kw)
    bs = [T.read_fastq(f) for f in d["fq"]]
    b = T.interleave(*bs) if paired else bs[0]
    codes, counts = T.load_table_arrays(d["tab"])
    return d, bs, b, codes, counts


@pytest.mark.parametrize("name,k,paired,maxcor,kw", CASES, ids=[c[0] for c in CASES])
def test_correct_batch_matches_oracle(gpu, workdir, name, k, paired, maxcor, kw):
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    mode = T.INTERLEAVED if paired else T.SINGLE
    exp = T.oracle_correct(table, b, k, mode, er, bq, maxcor)

    gpu.build_table(k, codes, counts)
    assert gpu.table_info()["entries"] == len(table)
    gpu.set_params(k, er, bq, maxcor)
    got = gpu.correct(mode, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["kernel_launches"] in (5, 6)

    names = ["seq", "correction", "l", "m", "h"]
    for nm, e, g in zip(names, exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d (exp %r got %r)" % (nm, len(bad), bad[0], e[bad[0]], g[bad[0]])
    assert int((exp[1] > 0).sum()) > 0  # the case really exercises corrections
    _check_compact(gpu, mode, b, bq, exp)


@pytest.mark.parametrize("name,k,paired,maxcor,kw", CASES[:4], ids=[c[0] for c in CASES[:4]])
def test_scan_matches_oracle(gpu, workdir, name, k, paired, maxcor, kw):
    """Stage parity for the lookup kernel: counts[] and GetStrongTrustedThreshold."""
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    b = b.slice(0, min(b.n, 4000))
    table = T.OracleTable(k)
    table.put(codes, counts)
    exp_counts, exp_thr = T.oracle_scan_counts(table, b)
    gpu.build_table(k, codes, counts)
    got_counts, got_thr = gpu.scan(b.seq, b.off)
    # only the first kcnt entries of each read are defined
    lens = np.diff(b.off.astype(np.int64))
    valid = np.zeros(len(b.seq), dtype=bool)
    for r in range(b.n):
        kc = max(0, lens[r] - k + 1)
        valid[b.off[r] : b.off[r] + kc] = True
    assert np.array_equal(exp_counts[valid], got_counts[valid])
    assert np.array_equal(exp_thr, got_thr)


def test_table_lookup_and_duplicates(gpu):
    """Store::Put semantics: canonicalisation, count<=1 skipped, last writer wins."""
    rng = np.random.default_rng(3)
    k = 23
    codes = rng.integers(0, 1 << 46, size=200000, dtype=np.uint64)
    counts = rng.integers(0, 500, size=len(codes), dtype=np.int32)
    # duplicates (same code again, later record) and reverse-complement duplicates
    codes = np.concatenate([codes, codes[:5000]])
    counts = np.concatenate([counts, rng.integers(2, 9, size=5000, dtype=np.int32)])
    table = T.OracleTable(k)
    table.put(codes, counts)
    gpu.build_table(k, codes, counts)
    assert gpu.table_info()["entries"] == len(table)
    q = np.concatenate([codes[:50000], rng.integers(0, 1 << 46, size=50000, dtype=np.uint64)])
    exp = np.array([T.oracle().orc_table_count(table.h, int(x)) for x in q], dtype=np.int32)
    assert np.array_equal(gpu.lookup(q), exp)


def test_edge_cases(gpu, workdir):
    """Empty batch, reads shorter than k, all-N, poly-A, exactly-k reads, N-rich reads."""
    k = 23
    d, bs, b, codes, counts = _load(workdir, "edge", k, False, dict(seed=21, ntx=10, reads=2000, len=80, err=0.02))
    table = T.OracleTable(k)
    table.put(codes, counts)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, 0.01, 48, 4)
    # empty batch
    so, corr, l, m, h = gpu.correct(T.SINGLE, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint32))
    assert len(corr) == 0
    base = b.seq[: b.off[1]].tobytes()
    reads = [b"ACGT", b"N" * 40, b"A" * 60, base[:k], base[: k + 1], base[:40] + b"NNNNNNN" + base[47:80],
             base[:30] + b"N" + base[31:80], b"T" * 22 + base[22:80], base[:79], b"ACGTN" * 12, base]
    seq = np.frombuffer(b"".join(reads), dtype=np.uint8)
    qual = np.full(len(seq), ord("I"), dtype=np.uint8)
    off = np.zeros(len(reads) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(r) for r in reads])
    eb = T.Batch([b"@x"] * len(reads), seq, qual, off)
    exp = T.oracle_correct(table, eb, k, T.SINGLE, 0.01, 48, 4)
    got = gpu.correct(T.SINGLE, eb.seq, eb.qual, eb.off)
    for e, g in zip(exp, got):
        assert np.array_equal(e, g)


def _code(s):
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.mark.parametrize("k", [2, 3, 15, 22, 23, 31, 32])
def test_context_table_semantics(gpu, k):
    """The table is keyed by (k-1)-mer contexts (two placements per k-mer): overlapping k-mers from both strands,
    reverse-complement duplicates (last record wins), palindromic contexts, counts beyond the 24-bit fields."""
    import random

    rnd = random.Random(100 + k)
    kmers = []
    for _ in range(60):  # chains of overlapping k-mers, random strand per k-mer
        s = "".join(rnd.choice("ACGT") for _ in range(k + 200))
        for j in range(len(s) - k + 1):
            x = s[j : j + k]
            kmers.append(x if rnd.random() < 0.5 else _rc(x))
    if k >= 3 and (k - 1) % 2 == 0:  # palindromic (k-1)-mer contexts extended on both sides
        for _ in range(200):
            half = "".join(rnd.choice("ACGT") for _ in range((k - 1) // 2))
            p = half + _rc(half)
            for b in "ACGT":
                kmers.append(p + b)
                kmers.append(b + p)
    kmers += ["A" * k, "T" * k, "AC" * (k // 2) + "A" * (k % 2), ("ACGT" * 9)[:k]]
    n0 = len(kmers)
    kmers += [_rc(x) for x in kmers[:: 7]]  # the other strand again, later in the file
    kmers += kmers[:: 11]                   # plain duplicates
    counts = [rnd.choice([0, 1, 2, 3, 50, 70000, (1 << 24) - 2, (1 << 24) - 1, 1 << 24, (1 << 31) - 1]) for _ in kmers]
    codes = np.array([_code(x) for x in kmers], dtype=np.uint64)
    counts = np.array(counts, dtype=np.int32)
    table = T.OracleTable(k)
    table.put(codes, counts)
    gpu.build_table(k, codes, counts)
    assert gpu.table_info()["entries"] == len(table)
    # queries: everything stored (both strands), one-base neighbours, random k-mers
    q = list(kmers[:n0]) + [_rc(x) for x in kmers[:n0]]
    for x in kmers[:2000]:
        i = rnd.randrange(k)
        q.append(x[:i] + rnd.choice("ACGT") + x[i + 1 :])
    q += ["".join(rnd.choice("ACGT") for _ in range(k)) for _ in range(2000)]
    qc = np.array([_code(x) for x in q], dtype=np.uint64)
    exp = np.array([T.oracle().orc_table_count(table.h, int(x)) for x in qc], dtype=np.int32)
    got = gpu.lookup(qc)
    bad = np.nonzero(exp != got)[0]
    assert len(bad) == 0, "%d mismatches, first %s exp %d got %d" % (len(bad), q[bad[0]], exp[bad[0]], got[bad[0]])


@pytest.mark.parametrize("chunk", [2, 998, 4096])
def test_chunked_host_batch_matches_oracle(gpu, workdir, chunk, monkeypatch):
    """rcx_correct_batch cuts a host batch into chunks that cycle through pipeline slots (copies overlap kernels);
    the result must not depend on the chunking (chunks start at arbitrary byte offsets, pairs stay together)."""
    name, k, paired, maxcor, kw = CASES[0]
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    if chunk == 2:
        b = b.slice(0, 600)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    exp = T.oracle_correct(table, b, k, T.INTERLEAVED, er, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, maxcor)
    monkeypatch.setenv("RCX_CHUNK_READS", str(chunk))
    got = gpu.correct(T.INTERLEAVED, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["kernel_launches"] >= 5 * ((b.n + chunk - 1) // chunk)
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        assert np.array_equal(np.asarray(e), np.asarray(g)), nm
    _check_compact(gpu, T.INTERLEAVED, b, bq, exp)  # the chunks' edits land in chunk order


def test_million_reads_full_pipeline(gpu, workdir):
    """A batch large enough to fill the persistent grids many times over and to span several pipeline chunks:
    1 M single-end reads against a 1.4 M-k-mer table, bit for bit against the oracle (8 host threads, ~10 s)."""
    k = 23
    d, bs, b, codes, counts = _load(workdir, "million", k, False, dict(seed=31, ntx=1000, reads=1000000, len=100, err=0.012, nrate=0.003))
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    exp = T.oracle_correct(table, b, k, T.SINGLE, er, bq, 4)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, 4)
    import os as _os
    _os.environ["RCX_CHUNK_READS"] = "262144"
    try:
        got = gpu.correct(T.SINGLE, b.seq, b.qual, b.off)
        _check_compact(gpu, T.SINGLE, b, bq, exp)
    finally:
        del _os.environ["RCX_CHUNK_READS"]
    st = gpu.stats()
    assert st["unsupported"] == 0 and st["work_items"] > 0 and st["second_pass"] > 0
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d" % (nm, len(bad), bad[0])


def test_table_build_in_partitions(gpu, monkeypatch):
    """'Last record wins' is resolved through a side array that covers one partition of the table at a time when HBM
    is short; force tiny partitions and compare with the one-partition build and the oracle."""
    rng = np.random.default_rng(5)
    k = 21
    base = rng.integers(0, 1 << 42, size=60000, dtype=np.uint64)
    codes = np.concatenate([base, base[:20000], base[5000:9000]])
    counts = rng.integers(0, 300, size=len(codes), dtype=np.int32)
    table = T.OracleTable(k)
    table.put(codes, counts)
    q = np.concatenate([base, rng.integers(0, 1 << 42, size=20000, dtype=np.uint64)])
    exp = np.array([T.oracle().orc_table_count(table.h, int(x)) for x in q], dtype=np.int32)
    gpu.build_table(k, codes, counts)
    one = gpu.lookup(q)
    monkeypatch.setenv("RCX_TABLE_WIN_SECTORS", "4096")
    gpu.build_table(k, codes, counts)
    assert gpu.table_info()["entries"] == len(table)
    parts = gpu.lookup(q)
    assert np.array_equal(one, exp) and np.array_equal(parts, exp)


def test_table_grows_and_densifies(gpu, monkeypatch):
    """Unrelated random k-mers have two (k-1)-mer contexts each, twice what the first sizing assumes: the build resizes;
    when the sparse size exceeds the memory budget it accepts a denser table (load up to 0.7) and still answers exactly."""
    rng = np.random.default_rng(9)
    k = 25
    codes = rng.integers(0, 1 << 50, size=300000, dtype=np.uint64)
    counts = rng.integers(2, 1000, size=len(codes), dtype=np.int32)
    table = T.OracleTable(k)
    table.put(codes, counts)
    q = np.concatenate([codes[:100000], rng.integers(0, 1 << 50, size=50000, dtype=np.uint64)])
    exp = np.array([T.oracle().orc_table_count(table.h, int(x)) for x in q], dtype=np.int32)
    gpu.build_table(k, codes, counts)
    sparse_bytes = gpu.table_info()["bytes"]
    assert np.array_equal(gpu.lookup(q), exp)
    monkeypatch.setenv("RCX_TABLE_MAX_BYTES", str(int(2 * len(codes) / 0.6 * 32)))   # room for load ~0.6 only
    gpu.build_table(k, codes, counts)
    assert gpu.table_info()["bytes"] < sparse_bytes
    assert gpu.table_info()["entries"] == len(table)
    assert np.array_equal(gpu.lookup(q), exp)


def test_long_reads_use_the_warp_kernels(gpu, workdir):
    """Reads longer than 255 bases do not fit the thread-per-read kernels' shared-memory rows: they take the
    warp-per-read scan / preparation / finalize kernels and the dense path[] / fix[] of the search kernel."""
    k = 27
    d, bs, b, codes, counts = _load(workdir, "long400", k, True, dict(seed=41, ntx=30, reads=3000, len=400, err=0.015, txmin=600, txmax=3000))
    assert int(np.diff(b.off.astype(np.int64)).max()) > 255
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    exp = T.oracle_correct(table, b, k, T.INTERLEAVED, er, bq, 4)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, 4)
    got = gpu.correct(T.INTERLEAVED, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["kernel_launches"] == 6 and st["unsupported"] == 0
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d" % (nm, len(bad), bad[0])
    assert int((exp[1] > 0).sum()) > 0
    _check_compact(gpu, T.INTERLEAVED, b, bq, exp)  # dense fix[] of the warp-per-read path


@pytest.mark.parametrize("name,k,paired,maxcor,kw", [CASES[0], CASES[3]], ids=[CASES[0][0], CASES[3][0]])
def test_warp_kernels_on_short_reads(gpu, workdir, name, k, paired, maxcor, kw, monkeypatch):
    """RCX_SCAN_WARP=1 sends ordinary reads through the warp-per-read kernels as well (same results)."""
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    mode = T.INTERLEAVED if paired else T.SINGLE
    exp = T.oracle_correct(table, b, k, mode, er, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, maxcor)
    monkeypatch.setenv("RCX_SCAN_WARP", "1")
    got = gpu.correct(mode, b.seq, b.qual, b.off)
    assert gpu.stats()["kernel_launches"] == 6
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        assert np.array_equal(np.asarray(e), np.asarray(g)), nm
    _check_compact(gpu, mode, b, bq, exp)


# Cases the GPU-vs-oracle campaign (tools/fuzz_gpu_vs_oracle.py, first run on hardware in round 2) found: small k with
# -maxcorK 8 puts more accepted edits on a read than the old byte-pair edit list of the search -> finalize hand-over held
# (rcx_stats.unsupported != 0 and differing reads).  The lists are 16-bit now and cannot overflow.
FUZZ_REGRESSIONS = [
    (23, False, 8, 0.5, dict(seed=934166292, ntx=20, reads=800, len=150, err=0.04, nrate=0.03, paralog=0.6, repeat=0.3, polya=0.3, short=0.0, sigma=2.0, txmin=300, txmax=1200)),
    (15, False, 8, 0.95, dict(seed=61331796, ntx=40, reads=2000, len=200, err=0.08, nrate=0.01, paralog=0.3, repeat=0.3, polya=0.3, short=0.02, sigma=2.0, txmin=300, txmax=1200)),
    (21, False, 8, 0.95, dict(seed=6873285, ntx=8, reads=2000, len=150, err=0.08, nrate=0.0, paralog=0.0, repeat=0.3, polya=0.9, short=0.02, sigma=2.0, txmin=300, txmax=1200)),
    (11, True, 8, 0.5, dict(seed=1014507394, ntx=20, reads=6000, len=100, err=0.003, nrate=0.03, paralog=0.0, repeat=0.3, polya=0.0, short=0.2, sigma=3.0, txmin=300, txmax=1200)),
    (17, True, 8, 0.95, dict(seed=696262160, ntx=40, reads=6000, len=125, err=0.01, nrate=0.03, paralog=0.0, repeat=0.6, polya=0.0, short=0.2, sigma=0.5, txmin=300, txmax=2500)),
]


@pytest.mark.parametrize("k,paired,maxcor,wk,kw", FUZZ_REGRESSIONS, ids=["k%d_%d" % (c[0], c[4]["len"]) for c in FUZZ_REGRESSIONS])
def test_many_accepted_edits_per_read(gpu, workdir, k, paired, maxcor, wk, kw):
    d, bs, b, codes, counts = _load(workdir, "fz%d_%d" % (k, kw["seed"] % 1000), k, paired, kw)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"], wk)
    bq = T.oracle_bad_quality(bs[0])
    mode = T.INTERLEAVED if paired else T.SINGLE
    exp = T.oracle_correct(table, b, k, mode, er, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, maxcor)
    got = gpu.correct(mode, b.seq, b.qual, b.off)
    assert gpu.stats()["unsupported"] == 0
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d" % (nm, len(bad), bad[0])
    _check_compact(gpu, mode, b, bq, exp)


def test_free_substitutions_spill_the_edit_row(gpu, workdir):
    """Substitutions on non-ACGT letters other than N are free (they do not count towards maxFixCnt) and the read filter
    only counts N, so one search path can carry far more edits than maxFixCnt: with -maxcorK 1 and k = 31 the shared-memory
    row of the search holds (250 / 31) * 2 + 9 = 25 of them and the rest goes to the spill area.  Reads peppered with 'R'."""
    k, maxcor = 31, 1
    d, bs, b, codes, counts = _load(workdir, "spill", k, False, dict(seed=77, ntx=12, reads=1500, len=250, err=0.0, nrate=0.0, txmin=600, txmax=2500, short=0.0))
    rng = np.random.default_rng(5)
    seq = b.seq.copy()
    for r in range(0, b.n, 3):  # every third read: an 'R' every ~6 bases beyond the first k
        o, e = int(b.off[r]), int(b.off[r + 1])
        pos = o + k + np.flatnonzero(rng.random(e - o - k) < 0.16)
        seq[pos] = ord("R")
    b = T.Batch(b.ids, seq, b.qual, b.off)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    exp = T.oracle_correct(table, b, k, T.SINGLE, er, bq, maxcor)
    assert int(exp[1].max()) > 30  # some read really gets more edits than the row holds
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, maxcor)
    got = gpu.correct(T.SINGLE, b.seq, b.qual, b.off)
    assert gpu.stats()["unsupported"] == 0
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d" % (nm, len(bad), bad[0])
    _check_compact(gpu, T.SINGLE, b, bq, exp)


def test_counts_of_2_to_the_29_take_the_warp_kernels(gpu, workdir):
    """A table with a count of 2^29 or more (poly-A in a very deep data set): the thread-per-read kernels keep 29 bits next to
    their flags, so the library routes such a table to the warp-per-read kernels (full 32-bit counts) at build time
    instead of flagging reads as unsupported."""
    name, k, paired, maxcor, kw = CASES[1]
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    counts = counts.copy()
    hot = np.argsort(-counts.astype(np.int64))[:50]
    counts[hot[:10]] = 600_000_000   # >= 2^29
    counts[hot[10:50]] = 30_000_000  # >= 2^24 - 1: the side list of saturated 24-bit fields
    table = T.OracleTable(k)
    table.put(codes, counts)
    bq = T.oracle_bad_quality(bs[0])
    exp = T.oracle_correct(table, b, k, T.SINGLE, 0.01, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, 0.01, bq, maxcor)
    got = gpu.correct(T.SINGLE, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["unsupported"] == 0 and st["kernel_launches"] == 6  # the warp-per-read kernels ran
    assert int(np.asarray(exp[4]).max()) >= 600_000_000           # the wide count reaches an h: tag
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        assert np.array_equal(np.asarray(e), np.asarray(g)), nm


@pytest.mark.parametrize("name,k,paired,maxcor,kw", [CASES[0], CASES[2], CASES[3], CASES[9]], ids=[CASES[0][0], CASES[2][0], CASES[3][0], CASES[9][0]])
def test_vh_mode_batches_the_very_heavy_items_first(gpu, workdir, name, k, paired, maxcor, kw, monkeypatch):
    """vh mode (picked by the library when a quarter of the previous batch's work items had nearly every k-window untrusted;
    forced here): the search kernel opens with warp batches over the listed very heavy items.  Only the order changes."""
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    table = T.OracleTable(k)
    table.put(codes, counts)
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    mode = T.INTERLEAVED if paired else T.SINGLE
    exp = T.oracle_correct(table, b, k, mode, er, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, er, bq, maxcor)
    monkeypatch.setenv("RCX_VH_MODE", "1")
    got = gpu.correct(mode, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["vheavy_items"] > 0 and st["unsupported"] == 0
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
        assert len(bad) == 0, "%s: %d mismatches, first at %d" % (nm, len(bad), bad[0])
    _check_compact(gpu, mode, b, bq, exp)


def test_large_counts_below_2_to_the_29_stay_on_the_thread_kernels(gpu, workdir):
    """Counts of millions (poly-A / repeat k-mers of a deep data set) on 150-bp reads: 3 M sits in the table's 24-bit fields,
    30 M in its side list of saturated fields; both stay below the 29 bits the thread-per-read kernels keep next to their
    flags, so those kernels run (5 launches) and the tags carry the large counts."""
    name, k, paired, maxcor, kw = [c for c in CASES if c[0] == "k31deep"][0]
    d, bs, b, codes, counts = _load(workdir, name, k, paired, kw)
    counts = counts.copy()
    hot = np.argsort(-counts.astype(np.int64))[:400]
    counts[hot[:200]] = 3_000_000
    counts[hot[200:400]] = 30_000_000
    table = T.OracleTable(k)
    table.put(codes, counts)
    bq = T.oracle_bad_quality(bs[0])
    mode = T.INTERLEAVED if paired else T.SINGLE
    exp = T.oracle_correct(table, b, k, mode, 0.01, bq, maxcor)
    gpu.build_table(k, codes, counts)
    gpu.set_params(k, 0.01, bq, maxcor)
    got = gpu.correct(mode, b.seq, b.qual, b.off)
    st = gpu.stats()
    assert st["unsupported"] == 0 and st["kernel_launches"] == 5
    assert int(np.asarray(exp[4]).max()) >= 3_000_000
    for nm, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
        assert np.array_equal(np.asarray(e), np.asarray(g)), nm
