"""CPU tests of the host-side I/O of the drop-in binary (rcorrector_b200/csrc/fastq_io.h) through the
small harness tests/harness/fastq_io_check.cpp (plain g++, no CUDA):
  * output file naming against what the REFERENCE binary really creates (Reads.h:39-75 GetFileName, :138-157);
  * the block line reader (plain, gzip, mapped-from-offset) against Python's own line splitting;
  * the parallel newline index; "%d" formatting; the gzip writer (level 1, like File.h's gzopen "w1")."""
import gzip
import os
import random
import shutil
import subprocess

import pytest

import rcx_testlib as T

HARNESS_SRC = os.path.join(T.ROOT, "tests", "harness", "fastq_io_check.cpp")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "fastq_io_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, HARNESS_SRC, "-lz"])

    def run(*args):
        return subprocess.run([exe] + [str(a) for a in args], check=True, capture_output=True, text=True).stdout.splitlines()

    return run


NAMES = ["reads.fq", "reads.fastq", "a.b.fq", "noext", "x.fastq.gz", "x.fq.gz", "x.fa.gz", "x.fasta.gz", "y.txt.gz", "plain.gz",
         "dot.dir/inner", "dot.dir/inner.fq", "two..dots.fq", ".hidden.fq", "trailing.", "UPPER.FQ", "r_1.fastq.gz.fq", "gz", "a.fqgz"]


@pytest.mark.skipif(not os.path.exists(T.REF_BIN), reason="oracle/_ref/rcorrector not built")
def test_output_names_match_the_reference_binary(harness, workdir):
    """Every name: run the reference on a tiny FASTQ stored under that name and see which file it creates."""
    d = T.synth(workdir, "names", k=15, paired=False, seed=3, ntx=3, reads=20, len=40)
    src = d["fq"][0]
    for i, name in enumerate(NAMES):
        base = os.path.join(workdir, "names_in_%d" % i)
        path = os.path.join(base, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if name.endswith("gz"):  # the reference opens anything ending in "gz" through zlib
            with open(src, "rb") as f, gzip.open(path, "wb") as g:
                g.write(f.read())
        else:
            shutil.copy(src, path)
        od = os.path.join(workdir, "names_out_%d" % i)
        os.makedirs(od)
        subprocess.run([T.REF_BIN, "-r", path, "-c", d["dump"], "-k", "15", "-od", od], check=True, capture_output=True)
        made = sorted(os.listdir(od))
        assert len(made) == 1, (name, made)
        ours = harness("name", od, 1, path)[0]
        assert ours == os.path.join(od, made[0]), (name, made[0], ours)


def test_fasta_names(harness):
    assert harness("name", "o", 0, "x.fa", "d/x.fasta.gz") == ["o/x.cor.fa", "o/x.cor.fa.gz"]


def _expected_lines(data):
    out = []
    pos = 0
    while pos < len(data):
        nl = data.find(b"\n", pos)
        if nl < 0:
            out.append("%d 0 %d" % (len(data) - pos, data[pos]))
            break
        out.append("%d 1 %d" % (nl - pos, data[pos] if nl > pos else -1))
        pos = nl + 1
    return out


@pytest.mark.parametrize("tail", ["newline", "no_newline", "empty_lines"])
def test_line_reader_plain_gzip_and_mapped(harness, workdir, tail):
    rng = random.Random(7)
    lines = [bytes(rng.choice(b"ACGTN@+!IJ") for _ in range(rng.choice([0, 1, 7, 100, 151, 5000]))) for _ in range(4000)]
    data = b"\n".join(lines) + (b"\n" if tail != "no_newline" else b"x")
    if tail == "empty_lines":
        data += b"\n\n\n"
    p = os.path.join(workdir, "lines_%s.txt" % tail)
    open(p, "wb").write(data)
    with gzip.open(p + ".gz", "wb", compresslevel=1) as g:
        g.write(data)
    exp = _expected_lines(data)
    assert harness("lines", p) == exp
    assert harness("lines", p + ".gz") == exp
    start = data.find(b"\n", len(data) // 3) + 1  # the parallel reader hands over to the serial one at a record start
    got = harness("memlines", p, start)
    assert got[:-1] == _expected_lines(data[start:]) and got[-1] == "tell %d" % len(data)


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_newline_index(harness, workdir, threads):
    rng = random.Random(11)
    data = bytes(rng.choice(b"AC\nGT\n@") for _ in range(200_003))
    p = os.path.join(workdir, "nl_%d.txt" % threads)
    open(p, "wb").write(data)
    for lo, hi in ((0, 0), (12345, 150_000), (199_990, 0)):
        end = hi or len(data)
        exp = [str(i - lo) for i in range(lo, end) if data[i] == 10]
        assert harness("newlines", p, lo, hi, threads) == exp


def test_format_int(harness):
    vals = [0, 1, -1, 9, 10, 99, 100, 16777215, 2147483647, -2147483648, -2147483647]
    assert harness("int", *vals) == [str(v) for v in vals]


def test_gzip_writer_round_trip(harness, workdir):
    src = os.path.join(workdir, "gzw.txt")
    open(src, "wb").write(b"@r1 l:1 m:2 h:3 cor\nACGT\n+\nIIII\n" * 5000)
    harness("gzwrite", src, os.path.join(workdir, "gzw.out.gz"))
    harness("gzwrite", src, os.path.join(workdir, "gzw.out.txt"))
    assert gzip.open(os.path.join(workdir, "gzw.out.gz"), "rb").read() == open(src, "rb").read()
    assert open(os.path.join(workdir, "gzw.out.txt"), "rb").read() == open(src, "rb").read()
