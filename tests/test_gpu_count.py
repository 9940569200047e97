"""GPU: SURVEY 8f N3 — the k-mer counting stages of run_rcorrector.pl (Jellyfish: canonical counts, k-mers seen at least twice)
on the device, checked against tools/rcx_kmercount (the exact CPU counter that stands in for Jellyfish in every fixture),
and the correction that follows against the reference's bytes."""
import os
import subprocess

import numpy as np
import pytest

import rcx_testlib as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "rcorrector_b200", "bin", "rcorrector")


def _canon(codes, k):
    """min(code, reverse complement) of 2-bit codes (A0 C1 G2 T3, first base in the highest bits)."""
    c = codes.astype(np.uint64)
    rc = np.zeros_like(c)
    x = c.copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
        x >>= np.uint64(2)
    return np.minimum(c, rc)


def _as_dict(codes, counts, k):
    cc = _canon(np.asarray(codes, dtype=np.uint64), k)
    keep = np.asarray(counts) >= 2
    return dict(zip(cc[keep].tolist(), np.asarray(counts)[keep].tolist()))


@pytest.mark.parametrize("name", ["sample_k23", "paired_k23", "single_k15_short", "paired_k31_maxcor8", "single_repeat_n"])
def test_counts_equal_the_cpu_counter_and_correction_follows(gpu, name, workdir):
    meta, fq, dump, expected = T.load_golden(name, workdir)
    k, paired = meta["k"], meta["paired"]
    bs = [T.read_fastq(f) for f in fq]
    want = _as_dict(*T.dump_to_arrays(dump, k), k)
    codes, counts, distinct, inst = gpu.count_kmers(k, [(b.seq, b.off) for b in bs])
    assert len(set(codes.tolist())) == len(codes) and distinct >= len(codes)
    assert 0 < inst <= sum(int(np.maximum(np.diff(b.off.astype(np.int64)) - k + 1, 0).sum()) for b in bs)
    assert _as_dict(codes, counts, k) == want
    assert gpu.table_info()["entries"] == len(want)
    # the table finish() built is the one the correction runs on: the reference's bytes (the fixtures whose ERROR_RATE does not
    # depend on the dump's record order)
    if meta["stderr"]["rate6"] == "0.010000":
        b = T.interleave(*bs) if paired else bs[0]
        gpu.set_params(k, 0.01, meta["stderr"]["bad_qual"], meta["maxcorK"])
        so, corr, l, m, h = gpu.correct(T.INTERLEAVED if paired else T.SINGLE, b.seq, b.qual, b.off)
        for i, exp in enumerate(expected):
            assert T.format_fastq(b, so, corr, l, m, h, rows=range(i, b.n, 2) if paired else None) == exp


def test_counting_table_grows_and_batches_add_up(gpu, workdir):
    """Several add() calls, a table that starts at its minimum size and grows by re-insertion: same counts as one call."""
    k = 21
    d = T.synth(workdir, "cnt_grow", k=k, paired=False, seed=91, ntx=400, reads=300000, len=100, err=0.02)
    b = T.read_fastq(d["fq"][0])
    want = _as_dict(*T.dump_to_arrays(d["dump"], k), k)
    parts = [b.slice(lo, min(b.n, lo + 70000)) for lo in range(0, b.n, 70000)]
    codes, counts, distinct, inst = gpu.count_kmers(k, [(p.seq, p.off) for p in parts], capacity_hint=0)
    assert distinct > (1 << 19)  # more distinct k-mers (errors make singletons) than half the initial table: it grew
    assert _as_dict(codes, counts, k) == want
    codes1, counts1, _, _ = gpu.count_kmers(k, [(b.seq, b.off)], capacity_hint=distinct)
    assert _as_dict(codes1, counts1, k) == want


def test_binary_counts_its_own_kmers(workdir):
    """`rcorrector -countkmers` (no -c dump): the binary counts the k-mers of its input files on the GPU, then corrects."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    od = os.path.join(workdir, "countkmers")
    r = subprocess.run([EXE, "-p"] + fq + ["-countkmers", "-k", str(meta["k"]), "-od", od, "-t", "4"], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    info = T.parse_ref_stderr(r.stderr.decode())
    assert info == meta["stderr"]  # Stored / rate / bad quality / Processed lines as with the dump of the exact counter
    for f, exp in zip(meta["outputs"], expected):
        assert open(os.path.join(od, f[:-3]), "rb").read() == exp
