"""CPU: the oracle (C restatement) against the golden fixtures produced by the unmodified reference."""
import pytest

import rcx_testlib as T


@pytest.mark.parametrize("name", T.golden_cases())
def test_oracle_reproduces_reference_bytes(name, workdir):
    meta, fq, dump, expected = T.load_golden(name, workdir)
    k, paired = meta["k"], meta["paired"]
    bs = [T.read_fastq(f) for f in fq]
    b = T.interleave(*bs) if paired else bs[0]
    table = T.OracleTable(k)
    stored = table.load_dump(dump)
    assert stored == meta["stderr"]["stored"]                      # "Stored %d kmers"
    er = table.error_rate(dump, meta["wk"])
    assert repr(er) == meta["oracle_error_rate"]
    assert "%.6f" % er == meta["stderr"]["rate6"]                  # "Weak kmer threshold rate: %lf"
    bq = T.oracle_bad_quality(bs[0])
    assert bq == meta["stderr"]["bad_qual"]                        # "Bad quality threshold is '%c'"
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.INTERLEAVED if paired else T.SINGLE, er, bq, meta["maxcorK"])
    assert b.n == meta["stderr"]["reads"]
    assert int(corr[corr > 0].sum()) == meta["stderr"]["bases"]   # "Corrected %d bases."
    for i, exp in enumerate(expected):
        rows = range(i, b.n, 2) if paired else None
        assert T.format_fastq(b, so, corr, l, m, h, rows=rows) == exp


def test_golden_fixtures_exist():
    assert len(T.golden_cases()) >= 6


def test_dump_arrays_match_oracle_loader(workdir):
    """The (code, count) arrays handed to rcx_table_build describe the same table as the text loader."""
    meta, fq, dump, _ = T.load_golden("paired_k23", workdir)
    k = meta["k"]
    a = T.OracleTable(k)
    a.load_dump(dump)
    codes, counts = T.dump_to_arrays(dump, k)
    b = T.OracleTable(k)
    b.put(codes, counts)
    assert len(a) == len(b)
    lib = T.oracle()
    for c in codes[:2000]:
        assert lib.orc_table_count(a.h, int(c)) == lib.orc_table_count(b.h, int(c))
