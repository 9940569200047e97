"""CPU: the oracle against the reference binary itself on fresh seeded inputs (only where oracle/_ref was
built, i.e. in the container that has /root/reference; the GPU box relies on the committed fixtures)."""
import os

import pytest

import rcx_testlib as T

needs_ref = pytest.mark.skipif(not os.path.exists(T.REF_BIN), reason="oracle/_ref/rcorrector not built")

CASES = [
    ("p100", 23, True, [], dict(seed=31, ntx=30, reads=6000, len=100, err=0.012)),
    ("s100n", 23, False, [], dict(seed=32, ntx=20, reads=8000, len=100, err=0.02, nrate=0.01, paralog=0.3, repeat=0.3)),
    ("k31", 31, True, ["-maxcorK", "8"], dict(seed=33, ntx=10, reads=2500, len=150, err=0.04)),
    ("maxcor2", 23, True, ["-maxcorK", "2"], dict(seed=34, ntx=12, reads=4000, len=100, err=0.03, paralog=0.5, repeat=0.5)),
]


@needs_ref
@pytest.mark.parametrize("name,k,paired,extra,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference(workdir, name, k, paired, extra, kw):
    d = T.synth(workdir, "ref_" + name, k=k, paired=paired, **kw)
    err, outs = T.run_reference(d["fq"], d["dump"], k, os.path.join(workdir, "refout_" + name), threads=4, extra=extra)
    info = T.parse_ref_stderr(err)
    bs = [T.read_fastq(f) for f in d["fq"]]
    b = T.interleave(*bs) if paired else bs[0]
    table = T.OracleTable(k)
    assert table.load_dump(d["dump"]) == info["stored"]
    er = table.error_rate(d["dump"])
    bq = T.oracle_bad_quality(bs[0])
    assert bq == info["bad_qual"] and "%.6f" % er == info["rate6"]
    maxcor = int(extra[extra.index("-maxcorK") + 1]) if "-maxcorK" in extra else 4
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.INTERLEAVED if paired else T.SINGLE, er, bq, maxcor)
    for i, f in enumerate(outs):
        rows = range(i, b.n, 2) if paired else None
        assert T.format_fastq(b, so, corr, l, m, h, rows=rows) == open(f, "rb").read()


@needs_ref
def test_bundled_sample_md5(workdir):
    """BASELINE.json configs[0]: the bundled paired sample, k=23 (md5s recorded in SURVEY §8c)."""
    s1, s2 = "/root/reference/Sample/sample_read1.fq", "/root/reference/Sample/sample_read2.fq"
    if not os.path.exists(s1):
        pytest.skip("reference sample not present")
    T.ensure_tools()
    dump = os.path.join(workdir, "sample.dump")
    T.run([T.KMERCOUNT, "-k", "23", "-o", dump, s1, s2])
    bs = [T.read_fastq(s1), T.read_fastq(s2)]
    b = T.interleave(*bs)
    table = T.OracleTable(23)
    assert table.load_dump(dump) == 253
    so, corr, l, m, h = T.oracle_correct(table, b, 23, T.INTERLEAVED, table.error_rate(dump), T.oracle_bad_quality(bs[0]), 4)
    assert int(corr[corr > 0].sum()) == 21 and b.n == 70
    assert T.md5(T.format_fastq(b, so, corr, l, m, h, rows=range(0, b.n, 2))) == "d67b929e0ce992bddd34525c78df8fe5"
    assert T.md5(T.format_fastq(b, so, corr, l, m, h, rows=range(1, b.n, 2))) == "40724361c10ba362ea15d1ae0b6b59dc"
