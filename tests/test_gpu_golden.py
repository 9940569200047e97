"""GPU: the CUDA path through the C ABI against the committed golden fixtures (bytes written by the unmodified
reference), and the drop-in `rcorrector` binary against the same fixtures (stderr lines included)."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import rcx_testlib as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "rcorrector_b200", "bin", "rcorrector")


@pytest.mark.parametrize("name", T.golden_cases())
def test_c_abi_reproduces_reference_bytes(gpu, name, workdir):
    meta, fq, dump, expected = T.load_golden(name, workdir)
    k, paired = meta["k"], meta["paired"]
    bs = [T.read_fastq(f) for f in fq]
    b = T.interleave(*bs) if paired else bs[0]
    codes, counts = T.dump_to_arrays(dump, k)
    gpu.build_table(k, codes, counts)
    er = gpu.estimate_error_rate(codes, meta["wk"])          # main.cpp:310-358 with GPU probes
    assert repr(er) == meta["oracle_error_rate"]
    gpu.set_params(k, er, meta["stderr"]["bad_qual"], meta["maxcorK"])
    so, corr, l, m, h = gpu.correct(T.INTERLEAVED if paired else T.SINGLE, b.seq, b.qual, b.off)
    assert int(corr[corr > 0].sum()) == meta["stderr"]["bases"]
    for i, exp in enumerate(expected):
        rows = range(i, b.n, 2) if paired else None
        assert T.format_fastq(b, so, corr, l, m, h, rows=rows) == exp


def _run(args, cwd=None):
    return subprocess.run([EXE] + args, capture_output=True, cwd=cwd)


@pytest.mark.parametrize("name", T.golden_cases())
def test_binary_is_a_drop_in(name, workdir):
    """Same argv grammar, same stderr lines, same output names and bytes as the reference binary."""
    meta, fq, dump, expected = T.load_golden(name, workdir)
    od = os.path.join(workdir, "bin_" + name)
    args = (["-p"] + fq if meta["paired"] else ["-r", fq[0]]) + ["-c", dump, "-k", str(meta["k"]), "-od", od, "-t", "4"] + meta["extra"]
    r = _run(args)
    assert r.returncode == 0, r.stderr.decode()
    info = T.parse_ref_stderr(r.stderr.decode())
    assert info == meta["stderr"]
    for f, exp in zip(meta["outputs"], expected):
        got = open(os.path.join(od, f[:-3]), "rb").read()
        assert got == exp, f


def test_binary_modes_match(workdir):
    """-i interleaved == -p outputs re-interleaved; -p -stdout == same stream; .gz in -> .gz out."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    k = str(meta["k"])
    w = os.path.join(workdir, "modes")
    os.makedirs(w, exist_ok=True)
    # interleaved file
    a = open(fq[0], "rb").read().split(b"\n")
    b = open(fq[1], "rb").read().split(b"\n")
    inter = []
    for i in range(0, len(a) - 1, 4):
        inter += a[i:i + 4] + b[i:i + 4]
    ipath = os.path.join(w, "inter.fq")
    open(ipath, "wb").write(b"\n".join(inter) + b"\n")
    e1 = expected[0].split(b"\n")
    e2 = expected[1].split(b"\n")
    exp_inter = []
    for i in range(0, len(e1) - 1, 4):
        exp_inter += e1[i:i + 4] + e2[i:i + 4]
    exp_inter = b"\n".join(exp_inter) + b"\n"
    r = _run(["-i", ipath, "-c", dump, "-k", k, "-od", os.path.join(w, "o1")])
    assert r.returncode == 0, r.stderr.decode()
    assert open(os.path.join(w, "o1", "inter.cor.fq"), "rb").read() == exp_inter
    r = _run(["-p", fq[0], fq[1], "-c", dump, "-k", k, "-stdout", "-od", os.path.join(w, "o2")])
    assert r.returncode == 0 and r.stdout == exp_inter
    # gz input -> gz output at the reference's name
    gz1 = os.path.join(w, "g_1.fq.gz")
    with open(fq[0], "rb") as f, gzip.open(gz1, "wb") as g:
        shutil.copyfileobj(f, g)
    r = _run(["-r", gz1, "-c", dump, "-k", k, "-od", os.path.join(w, "o3")])
    assert r.returncode == 0, r.stderr.decode()
    r2 = _run(["-r", fq[0], "-c", dump, "-k", k, "-od", os.path.join(w, "o4")])
    assert r2.returncode == 0
    assert gzip.open(os.path.join(w, "o3", "g_1.cor.fq.gz"), "rb").read() == open(os.path.join(w, "o4", "reads_1.cor.fq"), "rb").read()
    # two GPUs' worth of sharding logic on one device is exercised by -batch (many small device batches)
    r = _run(["-p", fq[0], fq[1], "-c", dump, "-k", k, "-od", os.path.join(w, "o5"), "-batch", "64"])
    assert r.returncode == 0
    assert open(os.path.join(w, "o5", "reads_1.cor.fq"), "rb").read() == expected[0]
    assert open(os.path.join(w, "o5", "reads_2.cor.fq"), "rb").read() == expected[1]


def test_perl_driver_stage3(workdir):
    """run_rcorrector.pl ... -stage 3 finds our binary next to it and produces the reference's bytes.
    The Perl driver is part of the reference: it is copied from /root/reference at test time, only where that exists."""
    src = "/root/reference/run_rcorrector.pl"
    if not os.path.exists(src) or shutil.which("perl") is None:
        pytest.skip("reference Perl driver not available on this box")
    import hashlib

    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    w = os.path.join(workdir, "perl")
    os.makedirs(os.path.join(w, "tmp"), exist_ok=True)
    shutil.copy(src, os.path.join(w, "run_rcorrector.pl"))
    shutil.copy(EXE, os.path.join(w, "rcorrector"))
    shutil.copy(os.path.join(ROOT, "rcorrector_b200", "librcx.so"), os.path.join(w, "librcx.so"))
    crc = hashlib.md5(("%s %s " % (fq[0], fq[1])).encode()).hexdigest()
    shutil.copy(dump, os.path.join(w, "tmp", "tmp_%s.jf_dump" % crc))
    env = dict(os.environ, LD_LIBRARY_PATH=w + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run(["perl", os.path.join(w, "run_rcorrector.pl"), "-1", fq[0], "-2", fq[1], "-k", str(meta["k"]), "-od",
                        os.path.join(w, "out"), "-tmpd", os.path.join(w, "tmp"), "-stage", "3"], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr.decode()
    assert open(os.path.join(w, "out", "reads_1.cor.fq"), "rb").read() == expected[0]
    assert open(os.path.join(w, "out", "reads_2.cor.fq"), "rb").read() == expected[1]


def _variant(workdir, name, data1, data2=None):
    w = os.path.join(workdir, "irr_" + name)
    os.makedirs(w, exist_ok=True)
    p1 = os.path.join(w, "a_1.fq")
    open(p1, "wb").write(data1)
    p2 = None
    if data2 is not None:
        p2 = os.path.join(w, "a_2.fq")
        open(p2, "wb").write(data2)
    return w, p1, p2


@pytest.mark.parametrize("case", ["no_final_newline", "ragged_quality", "crlf", "odd_interleaved", "plus_line_text", "regular_t1"])
def test_parallel_reader_equals_serial_reader(case, workdir):
    """The binary reads plain FASTQ with a parallel mapped-file reader and hands irregular input to the record-by-record
    reader that follows the reference's line semantics: both must give the same bytes (and the reference's, where the
    reference binary is on the box)."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    k = str(meta["k"])
    d1, d2 = open(fq[0], "rb").read(), open(fq[1], "rb").read()
    args_extra = []
    if case == "no_final_newline":
        d1, d2 = d1[:-1], d2[:-1]
    elif case == "ragged_quality":
        lines = d1.split(b"\n")
        lines[7] = lines[7][:-3]          # second record: quality shorter than the sequence
        lines[4 * 50 + 3] += b"II"        # another record: longer
        d1 = b"\n".join(lines)
    elif case == "crlf":
        d1, d2 = d1.replace(b"\n", b"\r\n"), d2.replace(b"\n", b"\r\n")
    elif case == "plus_line_text":
        lines = d1.split(b"\n")
        for i in range(2, len(lines) - 1, 4):
            lines[i] = b"+" + lines[i - 2][1:]
        d1 = b"\n".join(lines)
    mode = ["-p"]
    if case == "odd_interleaved":
        a, b = d1.split(b"\n"), d2.split(b"\n")
        inter = []
        for i in range(0, len(a) - 1, 4):
            inter += a[i:i + 4] + b[i:i + 4]
        d1, d2 = b"\n".join(inter[:-4]) + b"\n", None   # drop the last mate: odd number of records
        mode = ["-i"]
    w, p1, p2 = _variant(workdir, case, d1, d2)
    files = mode + [p1] + ([p2] if p2 else [])
    t = "1" if case == "regular_t1" else "4"
    outs = {}
    for tag, env in (("fast", {}), ("serial", {"RCX_SERIAL_IO": "1"})):
        od = os.path.join(w, "o_" + tag)
        r = subprocess.run([EXE] + files + ["-c", dump, "-k", k, "-od", od, "-t", t, "-batch", "512"], capture_output=True,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr.decode()
        outs[tag] = [open(os.path.join(od, f), "rb").read() for f in sorted(os.listdir(od))]
        outs[tag + "_err"] = T.parse_ref_stderr(r.stderr.decode())
    assert outs["fast"] == outs["serial"]
    assert outs["fast_err"] == outs["serial_err"]
    ref = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
    if os.path.exists(ref) and case in ("no_final_newline", "plus_line_text", "regular_t1"):
        od = os.path.join(w, "o_ref")
        os.makedirs(od, exist_ok=True)
        r = subprocess.run([ref] + files + ["-c", dump, "-k", k, "-od", od, "-t", "2"], capture_output=True)
        assert r.returncode == 0
        assert [open(os.path.join(od, f), "rb").read() for f in sorted(os.listdir(od))] == outs["fast"]


def test_dump_parsers_agree(workdir):
    """Parallel dump loader (regular '>count\\nKMER\\n' files) vs the serial token parser (anything else): same table, same
    ERROR_RATE, same output."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    k = str(meta["k"])
    w = os.path.join(workdir, "dumps")
    os.makedirs(w, exist_ok=True)
    spaced = os.path.join(w, "spaced.dump")   # same tokens, separated by blanks: only the token parser accepts it
    open(spaced, "wb").write(open(dump, "rb").read().replace(b"\n", b" ") + b"\n")
    res = []
    for tag, dp, env in (("par", dump, {}), ("ser", dump, {"RCX_SERIAL_IO": "1"}), ("spaced", spaced, {})):
        od = os.path.join(w, tag)
        r = subprocess.run([EXE, "-p", fq[0], fq[1], "-c", dp, "-k", k, "-od", od, "-t", "4"], capture_output=True,
                           env=dict(os.environ, RCX_DEBUG_SCALARS="1", **env))
        assert r.returncode == 0, r.stderr.decode()
        err = r.stderr.decode()
        res.append((T.parse_ref_stderr(err), [l for l in err.splitlines() if l.startswith("RCX_DEBUG")],
                    [open(os.path.join(od, f), "rb").read() for f in sorted(os.listdir(od))]))
    assert res[0] == res[1] == res[2]
    assert res[0][2] == expected


def _fasta_of(fq_path, out_path):
    lines = open(fq_path, "rb").read().split(b"\n")
    fa = []
    for i in range(0, len(lines) - 1, 4):
        fa += [b">" + lines[i][1:], lines[i + 1]]
    open(out_path, "wb").write(b"\n".join(fa) + b"\n")


@pytest.mark.parametrize("name", ["single_repeat_n", "paired_k23"])
def test_fasta_input_through_the_product_binary(name, workdir):
    """SURVEY §8f N4 on hardware: FASTA in -> <name>.cor.fa out.  Quality is absent (qual[0] == 0), so the finalize kernels
    take the `no quality` side of the three post-filters.  Compared with the unmodified reference (-t 2: its -t 1 FASTA path
    dereferences a NULL pointer) when it travelled to this box, and always with the CPU oracle fed zero qualities."""
    meta, fq, dump, expected = T.load_golden(name, workdir)
    w = os.path.join(workdir, "fasta_" + name)
    os.makedirs(w, exist_ok=True)
    fa = []
    for i, f in enumerate(fq):
        fa.append(os.path.join(w, "x_%d.fa" % (i + 1)))
        _fasta_of(f, fa[-1])
    k = meta["k"]
    files = ["-p"] + fa if meta["paired"] else ["-r", fa[0]]
    od = os.path.join(w, "ours")
    r = _run(files + ["-c", dump, "-k", str(k), "-od", od, "-t", "2"] + meta["extra"])
    assert r.returncode == 0, r.stderr.decode()
    got = [open(os.path.join(od, "x_%d.cor.fa" % (i + 1)), "rb").read() for i in range(len(fa))]
    # oracle: same reads with zero qualities and badQualityThreshold 0 (GetBadQuality returns 0 without quality)
    bs = [T.read_fastq(f) for f in fq]
    b = T.interleave(*bs) if meta["paired"] else bs[0]
    b.qual = np.zeros_like(b.qual)
    table = T.OracleTable(k)
    table.load_dump(dump)
    er = table.error_rate(dump, meta["wk"])
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.INTERLEAVED if meta["paired"] else T.SINGLE, er, 0, meta["maxcorK"])
    for i in range(len(fa)):
        rows = range(i, b.n, 2) if meta["paired"] else range(b.n)
        exp = b"".join(b">" + b.ids[x][1:] + b" l:%d m:%d h:%d" % (l[x], m[x], h[x]) +
                       (b"" if corr[x] == 0 else (b" unfixable_error" if corr[x] == -1 else b" cor")) + b"\n" +
                       so[b.off[x]:b.off[x + 1]].tobytes() + b"\n" for x in rows)
        assert got[i] == exp
    ref = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
    if os.path.exists(ref):
        odr = os.path.join(w, "ref")
        rr = subprocess.run([ref] + files + ["-c", dump, "-k", str(k), "-od", odr, "-t", "2"] + meta["extra"], capture_output=True)
        assert rr.returncode == 0
        assert [open(os.path.join(odr, "x_%d.cor.fa" % (i + 1)), "rb").read() for i in range(len(fa))] == got
        assert T.parse_ref_stderr(rr.stderr.decode()) == T.parse_ref_stderr(r.stderr.decode())


def test_fastq_then_fasta_on_one_command_line(workdir):
    """DESIGN §9, pinned: `-r a.fq -r b.fa`.  The quality threshold is the one the reference prints for this command line
    (GetBadQuality counts the last FASTQ record's quality line again for every FASTA record, reproduced); in the correction
    itself a FASTA record has NO quality here (zeros), whatever the batch slot held before it -- the reference's post-filters
    read those stale bytes, which is the one documented deviation.  Expected bytes: the CPU oracle with that threshold."""
    meta, fq, dump, expected = T.load_golden("single_repeat_n", workdir)
    w = os.path.join(workdir, "mixed")
    os.makedirs(w, exist_ok=True)
    fa = os.path.join(w, "b.fa")
    _fasta_of(fq[0], fa)
    k = meta["k"]
    args = ["-r", fq[0], "-r", fa, "-c", dump, "-k", str(k), "-t", "2"]
    r = _run(args + ["-od", os.path.join(w, "both")])
    assert r.returncode == 0, r.stderr.decode()
    info = T.parse_ref_stderr(r.stderr.decode("latin-1"))
    ref = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
    if os.path.exists(ref):
        rr = subprocess.run([ref] + args + ["-od", os.path.join(w, "ref")], capture_output=True)
        ri = T.parse_ref_stderr(rr.stderr.decode("latin-1"))
        assert (info["stored"], info["rate6"], info["bad_qual"], info["reads"]) == (ri["stored"], ri["rate6"], ri["bad_qual"], ri["reads"])
    bq = info["bad_qual"] - 256 if info["bad_qual"] >= 128 else info["bad_qual"]
    b = T.read_fastq(fq[0])
    table = T.OracleTable(k)
    table.load_dump(dump)
    er = table.error_rate(dump, meta["wk"])
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.SINGLE, er, bq, meta["maxcorK"])
    assert open(os.path.join(w, "both", "reads_1.cor.fq"), "rb").read() == T.format_fastq(b, so, corr, l, m, h)
    b0 = T.Batch(b.ids, b.seq, np.zeros_like(b.qual), b.off)
    so, corr, l, m, h = T.oracle_correct(table, b0, k, T.SINGLE, er, bq, meta["maxcorK"])
    exp = b"".join(b">" + b.ids[x][1:] + b" l:%d m:%d h:%d" % (l[x], m[x], h[x]) +
                   (b"" if corr[x] == 0 else (b" unfixable_error" if corr[x] == -1 else b" cor")) + b"\n" +
                   so[b.off[x]:b.off[x + 1]].tobytes() + b"\n" for x in range(b.n))
    assert open(os.path.join(w, "both", "b.cor.fa"), "rb").read() == exp


def test_gpus_2_shards_a_batch_over_two_devices(workdir):
    """-gpus 2: one host thread and one context per GPU, the table uploaded once and copied device to device, contiguous
    record shards (mates together), output in input order — the reference's bytes."""
    import ctypes

    try:
        n = ctypes.c_int(0)
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaGetDeviceCount(ctypes.byref(n))
        ndev = n.value
    except OSError:
        import torch
        ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    for name in ("paired_k23", "paired_rate_default", "single_wk01_rate"):
        meta, fq, dump, expected = T.load_golden(name, workdir)
        od = os.path.join(workdir, "g2_" + name)
        args = (["-p"] + fq if meta["paired"] else ["-r", fq[0]]) + ["-c", dump, "-k", str(meta["k"]), "-od", od, "-t", "4", "-gpus", "2"] + meta["extra"]
        r = _run(args)
        assert r.returncode == 0, r.stderr.decode()
        assert T.parse_ref_stderr(r.stderr.decode()) == meta["stderr"]
        for f, exp in zip(meta["outputs"], expected):
            assert open(os.path.join(od, f[:-3]), "rb").read() == exp, f


def test_unsupported_stop_leaves_no_partial_output(workdir):
    """rcx_stats.unsupported != 0 (a read the CUDA path cannot reproduce bit for bit; forced here through the library's test
    hook): the binary stops with an error, exit code 1, and removes the outputs it had started instead of leaving
    truncated *.cor.fq files behind."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir)
    od = os.path.join(workdir, "unsup")
    r = subprocess.run([EXE, "-p"] + fq + ["-c", dump, "-k", str(meta["k"]), "-od", od, "-batch", "64"], capture_output=True,
                       env=dict(os.environ, RCX_TEST_FORCE_UNSUPPORTED="3"))
    assert r.returncode == 1
    assert b"does not reproduce" in r.stderr
    assert not os.path.exists(od) or os.listdir(od) == []
