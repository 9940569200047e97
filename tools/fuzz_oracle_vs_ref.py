#!/usr/bin/env python
"""CPU differential campaign: the oracle (oracle/librcx_oracle.so) against the UNMODIFIED reference binary
(oracle/_ref/rcorrector) on random small workloads — random k, read length, error / N rates, repeats, paralogs,
-maxcorK, -wk, single / paired.  Test infrastructure only; needs /root/reference to have been compiled by
`make -C oracle ref` (this container).  Every case that differs is kept under --keep for a fixture.

usage: python tools/fuzz_oracle_vs_ref.py --cases 200 --seed 1 [--keep /tmp/fuzz_fail]
"""
import argparse
import os
import random
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcx_testlib as T  # noqa: E402


LONG = False
WILD_QUALS = False
LONG_LENGTHS = [256, 300, 400, 512, 640, 800, 1000]  # rcx_synth generates up to 1000 bases (the reference allows 1023)


def wild_quals(rng, files):
    """The generator's qualities stay inside a narrow band: rewrite part of them with the whole printable range."""
    for f in files:
        lines = open(f, "rb").read().split(b"\n")
        for i in range(3, len(lines), 4):
            r = rng.random()
            n = len(lines[i])
            if r < 0.25:
                lines[i] = bytes(rng.randrange(33, 127) for _ in range(n))
            elif r < 0.35:
                lines[i] = bytes(rng.choice(b"!\"#$%") for _ in range(n))
            elif r < 0.45:
                lines[i] = bytes(rng.choice(b"~}|IJ") for _ in range(n))
            elif r < 0.5 and n > 4:
                lines[i] = b"!" * 2 + lines[i][2:-2] + b"~" * 2
        open(f, "wb").write(b"\n".join(lines))


def one_case(rng, wd, idx):
    k = rng.choice([11, 13, 15, 17, 19, 21, 23, 23, 23, 25, 27, 29, 31, 31, 32])
    paired = rng.random() < 0.5
    length = rng.choice(LONG_LENGTHS if LONG else [k + 5, 50, 75, 100, 100, 125, 150, 150, 200, 250])
    if length < k + 2:
        length = k + 2
    kw = dict(seed=rng.randrange(1, 1 << 30), ntx=rng.choice([3, 8, 20, 40]), reads=rng.choice([100, 300] if LONG else [300, 800, 2000]), len=length,
              err=rng.choice([0.0, 0.003, 0.01, 0.02, 0.04, 0.08]), nrate=rng.choice([0.0, 0.002, 0.01, 0.03]),
              paralog=rng.choice([0.0, 0.0, 0.3, 0.6]), repeat=rng.choice([0.0, 0.0, 0.3, 0.6]), polya=rng.choice([0.0, 0.3, 0.9]),
              short=rng.choice([0.0, 0.02, 0.2]), sigma=rng.choice([0.5, 2.0, 3.0]), txmin=max(length, 300))
    kw["txmax"] = max(kw["txmin"], rng.choice([400, 1200, 2500]))
    extra = []
    maxcor = 4
    if rng.random() < 0.5:
        maxcor = rng.choice([1, 2, 3, 5, 6, 8])
        extra += ["-maxcorK", str(maxcor)]
    wk = 0.95
    if rng.random() < 0.25:
        wk = rng.choice([0.1, 0.5, 0.8, 0.99])
        extra += ["-wk", str(wk)]
    name = "f%04d" % idx
    d = T.synth(wd, name, k=k, paired=paired, **kw)
    if WILD_QUALS:
        wild_quals(rng, d["fq"])
    err, outs = T.run_reference(d["fq"], d["dump"], k, os.path.join(wd, "out_" + name), threads=rng.choice([1, 3]), extra=extra)
    info = T.parse_ref_stderr(err)
    bs = [T.read_fastq(f) for f in d["fq"]]
    b = T.interleave(*bs) if paired else bs[0]
    table = T.OracleTable(k)
    problems = []
    if table.load_dump(d["dump"]) != info["stored"]:
        problems.append("stored")
    er = table.error_rate(d["dump"], wk)
    bq = T.oracle_bad_quality(bs[0])
    if bq != info["bad_qual"]:
        problems.append("bad_qual %r vs %r" % (bq, info["bad_qual"]))
    if "%.6f" % er != info["rate6"]:
        problems.append("rate %.6f vs %s" % (er, info["rate6"]))
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.INTERLEAVED if paired else T.SINGLE, er, bq, maxcor)
    for i, f in enumerate(outs):
        rows = range(i, b.n, 2) if paired else None
        if T.format_fastq(b, so, corr, l, m, h, rows=rows) != open(f, "rb").read():
            problems.append("output %d" % i)
    desc = dict(k=k, paired=paired, extra=extra, **kw)
    stats = dict(reads=b.n, corrected=int(corr[corr > 0].sum()), unfixable=int((corr < 0).sum()))
    return problems, desc, stats, name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--long", action="store_true", help="reads of 256..1000 bases (the warp-per-read kernels' range)")
    ap.add_argument("--wild-quals", action="store_true", help="rewrite part of the quality strings with the whole printable range")
    a = ap.parse_args()
    global LONG, WILD_QUALS
    LONG = a.long
    WILD_QUALS = a.wild_quals
    if not os.path.exists(T.REF_BIN):
        sys.exit("oracle/_ref/rcorrector not built (make -C oracle ref)")
    rng = random.Random(a.seed)
    bad = 0
    tot = dict(reads=0, corrected=0, unfixable=0)
    t0 = time.time()
    for i in range(a.cases):
        wd = tempfile.mkdtemp(prefix="rcxfuzz_")
        try:
            problems, desc, stats, name = one_case(rng, wd, i)
            for k2 in tot:
                tot[k2] += stats[k2]
            if problems:
                bad += 1
                print("MISMATCH case %d: %s  %s" % (i, problems, desc), flush=True)
                if a.keep:
                    os.makedirs(a.keep, exist_ok=True)
                    shutil.copytree(wd, os.path.join(a.keep, name), dirs_exist_ok=True)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    print("fuzz seed %d: %d cases, %d mismatching, %d reads, %d corrected bases, %d unfixable reads, %.0f s" % (
        a.seed, a.cases, bad, tot["reads"], tot["corrected"], tot["unfixable"], time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
