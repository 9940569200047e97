#!/usr/bin/env python
"""GPU differential campaign (run on the B200 box through gpurun): the CUDA path through the C ABI against the CPU oracle on
random small workloads — the generator of tools/fuzz_oracle_vs_ref.py (k 11..32, 40..250 bp or --long 256..1000 bp, error /
N / repeat / paralog / poly-A rates, -maxcorK 1..8, single / paired), so that together the two campaigns tie
GPU == oracle == reference.  Compares corrected bases, correction, l, m, h per record and the ERROR_RATE estimate; any
`rcx_stats.unsupported` is reported too.  Test infrastructure (the oracle is the checker).

usage (GPU box): python tools/fuzz_gpu_vs_oracle.py --cases 200 --seed 1 [--long]
"""
import argparse
import os
import random
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_oracle_vs_ref as F  # noqa: E402
import rcx_testlib as T  # noqa: E402


def case_params(rng):
    k = rng.choice([11, 13, 15, 17, 19, 21, 23, 23, 23, 25, 27, 29, 31, 31, 32])
    paired = rng.random() < 0.5
    length = rng.choice(F.LONG_LENGTHS if F.LONG else [k + 5, 50, 75, 100, 100, 125, 150, 150, 200, 250])
    length = max(length, k + 2)
    kw = dict(seed=rng.randrange(1, 1 << 30), ntx=rng.choice([3, 8, 20, 40]), reads=rng.choice([100, 300] if F.LONG else [300, 800, 2000, 6000]),
              len=length, err=rng.choice([0.0, 0.003, 0.01, 0.02, 0.04, 0.08]), nrate=rng.choice([0.0, 0.002, 0.01, 0.03]),
              paralog=rng.choice([0.0, 0.0, 0.3, 0.6]), repeat=rng.choice([0.0, 0.0, 0.3, 0.6]), polya=rng.choice([0.0, 0.3, 0.9]),
              short=rng.choice([0.0, 0.02, 0.2]), sigma=rng.choice([0.5, 2.0, 3.0]), txmin=max(length, 300))
    kw["txmax"] = max(kw["txmin"], rng.choice([400, 1200, 2500]))
    maxcor = rng.choice([4, 4, 4, 1, 2, 3, 5, 6, 8])
    wk = rng.choice([0.95, 0.95, 0.95, 0.1, 0.5, 0.99])
    return k, paired, kw, maxcor, wk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--long", action="store_true")
    ap.add_argument("--wild-quals", action="store_true")
    ap.add_argument("--iupac", action="store_true", help="pepper part of the reads with non-ACGT letters other than N (free substitutions: long edit lists)")
    ap.add_argument("--chunks", action="store_true", help="random tiny RCX_CHUNK_READS per case (chunk boundaries, edit order of the compact interface)")
    ap.add_argument("--count", action="store_true", help="also count the k-mers of the reads on the GPU and compare with the dump")
    a = ap.parse_args()
    F.LONG = a.long
    from rcorrector_b200 import Corrector

    gpu = Corrector(0)  # raises without a CUDA device: there is no CPU path to compare
    rng = random.Random(a.seed)
    bad = 0
    tot = dict(reads=0, corrected=0, unsupported=0)
    t0 = time.time()
    for i in range(a.cases):
        k, paired, kw, maxcor, wk = case_params(rng)
        wd = tempfile.mkdtemp(prefix="rcxgpufuzz_")
        try:
            d = T.synth(wd, "g%04d" % i, k=k, paired=paired, **kw)
            if a.wild_quals:
                F.wild_quals(rng, d["fq"])
            bs = [T.read_fastq(f) for f in d["fq"]]
            if a.iupac:  # (after the dump was counted: the table stays that of the clean reads)
                nrng = np.random.default_rng(rng.randrange(1 << 30))
                for x in bs:
                    seq = x.seq.copy()
                    for r in range(0, x.n, rng.choice([2, 3, 5])):
                        o, e = int(x.off[r]), int(x.off[r + 1])
                        hit = o + np.flatnonzero(nrng.random(e - o) < rng.choice([0.02, 0.1, 0.25]))
                        seq[hit] = nrng.choice(np.frombuffer(b"RYKMSWBDHV", dtype=np.uint8), size=len(hit))
                    x.seq = seq
            b = T.interleave(*bs) if paired else bs[0]
            if a.chunks:
                os.environ["RCX_CHUNK_READS"] = str(rng.choice([26, 150, 334, 1000, 4000]))
            codes, counts = T.dump_to_arrays(d["dump"], k)
            table = T.OracleTable(k)
            table.load_dump(d["dump"])
            er = table.error_rate(d["dump"], wk)
            bq = T.oracle_bad_quality(bs[0])
            mode = T.INTERLEAVED if paired else T.SINGLE
            exp = T.oracle_correct(table, b, k, mode, er, bq, maxcor)
            gpu.build_table(k, codes, counts)
            er_gpu = gpu.estimate_error_rate(codes, wk)
            gpu.set_params(k, er, bq, maxcor)
            got = gpu.correct(mode, b.seq, b.qual, b.off)
            st = gpu.stats()
            problems = []
            # the compact host interface (quality bits in, edits out) on the same case
            from rcorrector_b200.api import apply_edits, quality_bits
            c_corr, c_l, c_m, c_h, c_edits = gpu.correct_compact(mode, b.seq, quality_bits(b.qual, bq), b.off)
            st["unsupported"] += gpu.stats()["unsupported"]
            got_c = (apply_edits(b.seq, b.off, c_corr, c_edits), c_corr, c_l, c_m, c_h)
            for name, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got_c):
                if not np.array_equal(np.asarray(e), np.asarray(g)):
                    problems.append("compact " + name)
            if repr(er_gpu) != repr(er):
                problems.append("error rate %r vs %r" % (er_gpu, er))
            for name, e, g in zip(["seq", "correction", "l", "m", "h"], exp, got):
                if not np.array_equal(np.asarray(e), np.asarray(g)):
                    problems.append(name)
            if st["unsupported"]:
                problems.append("unsupported=%d" % st["unsupported"])
            if a.count and not a.iupac:  # the dump is the exact count of these reads
                ccodes, ccounts, _, _ = gpu.count_kmers(k, [(x.seq, x.off) for x in bs])
                want = dict(zip(codes.tolist(), counts.tolist()))
                if dict(zip(ccodes.tolist(), ccounts.tolist())) != {kk: v for kk, v in want.items() if v >= 2}:
                    problems.append("k-mer counts")
            tot["reads"] += b.n
            tot["corrected"] += int(exp[1][exp[1] > 0].sum())
            tot["unsupported"] += int(st["unsupported"])
            if problems:
                bad += 1
                print("MISMATCH case %d: %s  k=%d paired=%s maxcorK=%d wk=%g %s" % (i, problems, k, paired, maxcor, wk, kw), flush=True)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    print("gpu fuzz seed %d%s: %d cases, %d mismatching, %d reads, %d corrected bases, %d unsupported, %.0f s" % (
        a.seed, " (long)" if a.long else "", a.cases, bad, tot["reads"], tot["corrected"], tot["unsupported"], time.time() - t0))
    gpu.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
