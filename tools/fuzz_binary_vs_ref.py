#!/usr/bin/env python
"""CPU differential campaign for the HOST side of the drop-in binary: rcorrector_main.cpp linked against the librcx test
double (tests/harness/rcx_double.cpp, the oracle as the device) against the UNMODIFIED reference binary, on random legal
FASTQ / FASTA inputs with awkward but well-formed content — ids with blanks, tabs and high bytes, '+' lines repeating the
id, quality strings that start with '@' or '+', reads shorter than k, of length 1 and up to 1023, CRLF line ends, a missing
final newline, single / paired / interleaved / several files, plain and gzip, -stdout, random -t and -batch.
Compares return code, stderr, stdout, file names and (decompressed) contents.  Test infrastructure only.

usage: python tools/fuzz_binary_vs_ref.py --cases 300 --seed 1
"""
import argparse
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcx_testlib as T  # noqa: E402


CRLF = False


def build_double(d):
    T.oracle()
    exe = os.path.join(d, "rcorrector")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "rcorrector_b200", "csrc", "rcorrector_main.cpp"),
                           os.path.join(ROOT, "tests", "harness", "rcx_double.cpp"), "-L" + os.path.join(ROOT, "oracle"), "-l:librcx_oracle.so",
                           "-lz", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def load_pairs(fq):
    """[(mate1, mate2)] with mate = (id, seq, qual) bytes, from the two files of a paired fixture."""
    bs = [T.read_fastq(f) for f in fq]
    rec = lambda b, i: (b.ids[i], b.seq[b.off[i]:b.off[i + 1]].tobytes(), b.qual[b.off[i]:b.off[i + 1]].tobytes())
    return [(rec(bs[0], i), rec(bs[1], i)) for i in range(bs[0].n)]


def mutate_records(rng, recs, style):
    """recs: list of (id, seq, qual) bytes.  Returns file bytes."""
    out = []
    eol = b"\r\n" if style["crlf"] else b"\n"
    for i, (rid, seq, qual) in enumerate(recs):
        if style["fasta"]:
            out.append(b">" + rid[1:] + eol + seq + eol)
            continue
        plus = b"+" + (rid[1:] if style["plus_id"] else b"")
        out.append(rid + eol + seq + eol + plus + eol + qual + eol)
    data = b"".join(out)
    if style["no_final_newline"] and data.endswith(eol):
        data = data[: -len(eol)]
    return data


def random_id(rng, i, mate):
    base = b"@r%d" % i
    extra = rng.choice([b"", b" desc", b"\tTAB", b"  two  blanks ", b" 1:N:0:ACGT", b" \xc3\xa9\xff", b"/%d" % (mate + 1), b" " + b"x" * rng.choice([1, 200, 1500])])
    return base + extra


def one_case(rng, exe, base_recs, dump, k, wd):
    n = rng.choice([1, 2, 3, 17, 100, 257, 600])
    paired = rng.random() < 0.4
    inter = (not paired) and rng.random() < 0.25
    # CRLF is off by default: the reference indexes nucToNum['\r' - 'A'] out of bounds (whatever lies before the table decides
    # what the base is); we treat every non-letter as non-ACGT
    style = dict(crlf=CRLF and rng.random() < 0.1, plus_id=rng.random() < 0.3, no_final_newline=rng.random() < 0.2, fasta=rng.random() < 0.12)
    gz = rng.random() < 0.3
    nfiles = rng.choice([1, 1, 1, 2])
    args, files = [], []
    first_kind = (paired, inter)
    for fno in range(nfiles):
        if fno > 0 and rng.random() < 0.5:  # a command line may mix -r, -p and -i files
            paired = rng.random() < 0.4
            inter = (not paired) and rng.random() < 0.4
        mates = []
        cnt = n if not inter else n - (n % 2)
        cnt = max(cnt, 2 if inter else 1)
        pick = [rng.randrange(len(base_recs)) for _ in range(cnt)]  # real mate pairs, so that the pair threshold matters
        for mate in range(2 if paired else 1):
            recs = []
            for i in range(cnt):
                rid, seq, qual = base_recs[pick[i // 2]][i % 2] if inter else base_recs[pick[i]][mate]
                L = rng.choice([len(seq), len(seq), len(seq), 1, 5, k - 1, k, k + 1, 60])
                L = min(L, len(seq))
                s, q = seq[:L], qual[:L]
                if rng.random() < 0.05:
                    q = bytes(rng.choice(b"@+!~IJ#5") for _ in range(L))
                if rng.random() < 0.02 and L > 4:
                    s = s[:2] + b"N" + s[3:]
                recs.append((random_id(rng, i, mate), s, q))
            if rng.random() < 0.05 and not style["fasta"]:  # one long read
                big = (base_recs[0][0][1] * 12)[: rng.choice([512, 1000, 1022])]  # 1022: the longest line the reference's 1024-byte fgets buffer takes whole
                recs[0] = (recs[0][0], big, b"I" * len(big))
            data = mutate_records(rng, recs, style)
            ext = (".fa" if style["fasta"] else rng.choice([".fq", ".fastq"])) + (".gz" if gz else "")
            path = os.path.join(wd, "f%d_%d%s" % (fno, mate + 1, ext))
            with open(path, "wb") as f:
                f.write(gzip.compress(data, 1) if gz else data)
            mates.append(path)
        args += (["-p"] + mates) if paired else (["-i" if inter else "-r", mates[0]])
    t = rng.choice(["1", "2", "3", "5"])
    if style["fasta"] and t == "1":
        t = "2"  # the reference's FASTA path crashes with -t 1
    paired, inter = first_kind
    common = ["-c", dump, "-k", str(k), "-t", t]
    if rng.random() < 0.2:
        common += ["-stdout"]
    if rng.random() < 0.2:
        common += ["-maxcorK", str(rng.choice([1, 2, 3, 6, 8]))]
    if rng.random() < 0.1:
        common += ["-wk", str(rng.choice([0.1, 0.5, 0.99]))]
    res = []
    for tag, binary, extra in (("ref", T.REF_BIN, []), ("ours", exe, ["-batch", str(rng.choice([2, 8, 64, 1000, 1 << 20]))])):
        od = os.path.join(wd, "out_" + tag)
        r = subprocess.run([binary] + args + common + ["-od", od] + extra, capture_output=True, cwd=wd)
        outs = {}
        if os.path.isdir(od):
            for name in sorted(os.listdir(od)):
                data = open(os.path.join(od, name), "rb").read()
                outs[name] = gzip.decompress(data) if name.endswith(".gz") and data else data
        res.append((r.returncode, r.stderr.replace(od.encode(), b"@OD"), r.stdout, outs))
    same = res[0] == res[1]
    desc = dict(n=n, paired=paired, inter=inter, gz=gz, nfiles=nfiles, t=t, args=args + common, **style)
    return same, desc, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--crlf", action="store_true", help="also generate CRLF files (known deviation)")
    a = ap.parse_args()
    global CRLF
    CRLF = a.crlf
    if not os.path.exists(T.REF_BIN):
        sys.exit("oracle/_ref/rcorrector not built (make -C oracle ref)")
    top = tempfile.mkdtemp(prefix="rcxbinfuzz_")
    exe = build_double(top)
    meta, fq, dump, _ = T.load_golden("paired_k23", top)
    k = meta["k"]
    base_recs = load_pairs(fq)
    rng = random.Random(a.seed)
    bad = 0
    for i in range(a.cases):
        wd = os.path.join(top, "c%04d" % i)
        os.makedirs(wd)
        same, desc, res = one_case(rng, exe, base_recs, dump, k, wd)
        if not same:
            bad += 1
            print("MISMATCH case %d: %s" % (i, desc), flush=True)
            for j, nm in enumerate(("rc", "stderr", "stdout")):
                if res[0][j] != res[1][j]:
                    print("   %s ref : %r" % (nm, res[0][j][:300] if j else res[0][j]))
                    print("   %s ours: %r" % (nm, res[1][j][:300] if j else res[1][j]))
            if res[0][3] != res[1][3]:
                print("   files ref : %s" % {kk: len(v) for kk, v in res[0][3].items()})
                print("   files ours: %s" % {kk: len(v) for kk, v in res[1][3].items()})
            if a.keep:
                shutil.copytree(wd, os.path.join(a.keep, "c%04d" % i), dirs_exist_ok=True)
        shutil.rmtree(wd, ignore_errors=True)
    print("binary fuzz seed %d: %d cases, %d mismatching" % (a.seed, a.cases, bad))
    shutil.rmtree(top, ignore_errors=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
