#!/usr/bin/env python
"""Debug aid: compact vs regular host interface at several chunk sizes on one seeded case (prints the first mismatches)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcx_testlib as T
from rcorrector_b200 import Corrector
from rcorrector_b200.api import apply_edits, quality_bits

wd = tempfile.mkdtemp()
k = 23
d = T.synth(wd, "p100", k=k, paired=True, seed=1, ntx=50, reads=20000, len=100, err=0.01)
bs = [T.read_fastq(f) for f in d["fq"]]
b = T.interleave(*bs).slice(0, 600)
codes, counts = T.load_table_arrays(d["tab"])
table = T.OracleTable(k); table.put(codes, counts)
er = table.error_rate(d["dump"]); bq = T.oracle_bad_quality(bs[0])
exp = T.oracle_correct(table, b, k, T.INTERLEAVED, er, bq, 4)
gpu = Corrector(0); gpu.build_table(k, codes, counts); gpu.set_params(k, er, bq, 4)
for chunk in (2, 4, 6, 16, 64, 600):
    os.environ["RCX_CHUNK_READS"] = str(chunk)
    for rep in range(3):
        got = gpu.correct(T.INTERLEAVED, b.seq, b.qual, b.off)
        corr, l, m, h, edits = gpu.correct_compact(T.INTERLEAVED, b.seq, quality_bits(b.qual, bq), b.off)
        st = gpu.stats()
        msg = []
        for nm, e, g in zip(["seq", "corr", "l", "m", "h"], exp, got):
            if not np.array_equal(np.asarray(e), np.asarray(g)): msg.append("regular " + nm)
        ne = int(np.maximum(corr, 0).sum())
        if len(edits) != ne: msg.append("edits %d vs sum %d" % (len(edits), ne))
        else:
            so = apply_edits(b.seq, b.off, corr, edits)
            for nm, e, g in zip(["seq", "corr", "l", "m", "h"], exp, (so, corr, l, m, h)):
                bad = np.nonzero(np.asarray(e) != np.asarray(g))[0]
                if len(bad): msg.append("compact %s: %d bad, first %d exp %r got %r" % (nm, len(bad), bad[0], np.asarray(e)[bad[0]], np.asarray(g)[bad[0]]))
        print("chunk", chunk, "rep", rep, "launches", st["kernel_launches"], "edits", len(edits), "->", msg or "ok", flush=True)
