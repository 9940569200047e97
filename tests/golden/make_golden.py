#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/rcorrector, built from /root/reference by oracle/Makefile) on small seeded inputs.

The reference ships no golden vectors of its own, so these files pin the oracle and the
CUDA path to the reference's behaviour.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Each fixture directory holds the inputs (FASTQ + Jellyfish-style dump, gzip'ed), the reference's
*.cor.fq outputs and meta.json (arguments and the scalars the reference prints on stderr).
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rcx_testlib as T  # noqa: E402

CASES = [
    # name, k, paired, reference extra args, synth kwargs
    ("paired_k23", 23, True, [], dict(seed=101, ntx=6, reads=500, len=100, err=0.02)),
    ("single_repeat_n", 23, False, [], dict(seed=102, ntx=5, reads=800, len=100, err=0.02, paralog=0.5, repeat=0.6, nrate=0.01)),
    ("paired_k31_maxcor8", 31, True, ["-maxcorK", "8"], dict(seed=103, ntx=4, reads=300, len=150, err=0.04)),
    ("single_wk01_rate", 23, False, ["-wk", "0.1"], dict(seed=104, ntx=1, reads=8000, len=100, err=0.01, txmin=300, txmax=350, sigma=0.3, chimera=0)),
    ("paired_rate_default", 23, True, [], dict(seed=107, ntx=1, reads=5000, len=100, err=0.015, txmin=320, txmax=360, chimera=0)),
    ("paired_maxcor1", 23, True, ["-maxcorK", "1"], dict(seed=105, ntx=5, reads=400, len=100, err=0.03, paralog=0.5, repeat=0.5)),
    ("single_k15_short", 15, False, [], dict(seed=106, ntx=6, reads=600, len=60, err=0.02, txmin=100)),
]


def gz(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


# BASELINE.json configs[0]: the reference's bundled paired sample (data files, not source), k = 23, with the exact-count
# dump Jellyfish would give (tools/rcx_kmercount).  md5 of the outputs: SURVEY §8c / BASELINE.md §2.
SAMPLE = ("sample_k23", 23, True, [], None)
SAMPLE_FILES = ["/root/reference/Sample/sample_read1.fq", "/root/reference/Sample/sample_read2.fq"]
SAMPLE_MD5 = ["d67b929e0ce992bddd34525c78df8fe5", "40724361c10ba362ea15d1ae0b6b59dc"]


def sample_inputs(wd):
    T.ensure_tools()
    fq = []
    for i, src in enumerate(SAMPLE_FILES):
        dst = os.path.join(wd, "reads_%d.fq" % (i + 1))  # the fixture loader expects reads_*.fq / reads.dump
        shutil.copy(src, dst)
        fq.append(dst)
    dump = os.path.join(wd, "reads.dump")
    T.run([T.KMERCOUNT, "-k", "23", "-o", dump] + fq)
    return {"fq": fq, "dump": dump}


def main():
    if not os.path.exists(T.REF_BIN):
        raise SystemExit("build oracle/_ref first (make -C oracle ref; needs /root/reference)")
    only = sys.argv[1:]
    for name, k, paired, extra, kw in CASES + [SAMPLE]:
        if only and name not in only:
            continue
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        with tempfile.TemporaryDirectory() as wd:
            d = sample_inputs(wd) if kw is None else T.synth(wd, "reads", k=k, paired=paired, **kw)
            err, outs = T.run_reference(d["fq"], d["dump"], k, os.path.join(wd, "out"), threads=1, extra=extra)
            err4, outs4 = T.run_reference(d["fq"], d["dump"], k, os.path.join(wd, "out4"), threads=4, extra=extra)
            for a, b in zip(outs, outs4):
                assert open(a, "rb").read() == open(b, "rb").read(), "-t 1 and -t 4 differ"
            info = T.parse_ref_stderr(err)
            if kw is None:
                assert [T.md5(open(f, "rb").read()) for f in outs] == SAMPLE_MD5, "bundled sample: md5 differs from SURVEY 8c"
            for f in d["fq"]:
                gz(f, os.path.join(out, os.path.basename(f) + ".gz"))
            gz(d["dump"], os.path.join(out, "reads.dump.gz"))
            for f in outs:
                gz(f, os.path.join(out, os.path.basename(f) + ".gz"))
            # the oracle's own ERROR_RATE with full precision (the reference prints 6 decimals only)
            t = T.OracleTable(k)
            t.load_dump(d["dump"])
            wk = float(extra[extra.index("-wk") + 1]) if "-wk" in extra else 0.95
            meta = {"k": k, "paired": paired, "extra": extra, "synth": kw, "stderr": info,
                    "oracle_error_rate": repr(t.error_rate(d["dump"], wk)), "wk": wk,
                    "maxcorK": int(extra[extra.index("-maxcorK") + 1]) if "-maxcorK" in extra else 4,
                    "inputs": [os.path.basename(f) + ".gz" for f in d["fq"]],
                    "outputs": [os.path.basename(f) + ".gz" for f in outs]}
            json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1, sort_keys=True)
            print(name, info)


if __name__ == "__main__":
    main()
