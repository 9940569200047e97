"""CPU: the HOST side of the drop-in `rcorrector` binary (argv grammar, dump ingest, badQualityThreshold, output naming,
parallel / serial FASTQ readers, the three-stage pipeline, per-GPU sharding of a batch, formatter, gzip writer) without a
GPU.  rcorrector_b200/csrc/rcorrector_main.cpp is compiled here against tests/harness/rcx_double.cpp — a TEST DOUBLE of
librcx.so whose "device" is the CPU oracle — into a temporary directory, and driven through the very same cases the GPU
suite runs against the product binary (tests/test_gpu_golden.py), plus sharding / batching cases.  Nothing of this is
shipped: the product binary links librcx.so only and fails without a CUDA device."""
import os
import subprocess

import pytest

import rcx_testlib as T
import test_gpu_golden as G

ROOT = T.ROOT
CSRC = os.path.join(ROOT, "rcorrector_b200", "csrc")
ORACLE_DIR = os.path.join(ROOT, "oracle")


@pytest.fixture(scope="module")
def double_exe(tmp_path_factory):
    T.oracle()  # builds oracle/librcx_oracle.so if needed
    d = tmp_path_factory.mktemp("double")
    exe = str(d / "rcorrector")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(CSRC, "rcorrector_main.cpp"),
                           os.path.join(ROOT, "tests", "harness", "rcx_double.cpp"), "-L" + ORACLE_DIR, "-l:librcx_oracle.so", "-lz",
                           "-Wl,-rpath," + ORACLE_DIR])
    return exe


@pytest.fixture(autouse=True)
def _use_double(monkeypatch, double_exe):
    monkeypatch.setattr(G, "EXE", double_exe)  # the shared cases call G._run / G.EXE


@pytest.mark.parametrize("name", T.golden_cases())
def test_host_side_reproduces_reference_bytes(name, workdir):
    G.test_binary_is_a_drop_in(name, workdir + "_cpu")


@pytest.mark.parametrize("name", ["single_repeat_n", "paired_k23"])
def test_fasta_through_the_binary(name, workdir):
    G.test_fasta_input_through_the_product_binary(name, workdir + "_cpu")


def test_fastq_then_fasta(workdir):
    G.test_fastq_then_fasta_on_one_command_line(workdir + "_cpu")


def test_countkmers_through_the_binary(workdir):
    """-countkmers: the binary's own pass over the input files (sequence lines of every file, mates included) feeds the
    counter; with the exact counts the rest of the run is the fixture's."""
    import test_gpu_count as GC

    GC.EXE = G.EXE
    GC.test_binary_counts_its_own_kmers(workdir + "_cpu")


def test_modes(workdir):
    G.test_binary_modes_match(workdir + "_cpu")


@pytest.mark.parametrize("case", ["no_final_newline", "ragged_quality", "crlf", "odd_interleaved", "plus_line_text", "regular_t1"])
def test_readers(case, workdir):
    G.test_parallel_reader_equals_serial_reader(case, workdir + "_cpu")


def test_dump_parsers(workdir):
    G.test_dump_parsers_agree(workdir + "_cpu")


def test_perl_driver_stage3(workdir):
    """run_rcorrector.pl -stage 3 (copied from /root/reference at test time, where that exists) drives the binary unchanged."""
    G.test_perl_driver_stage3(workdir + "_cpu")


@pytest.mark.parametrize("gpus,batch,threads", [(2, 64, 1), (3, 50, 5), (5, 1 << 20, 2), (8, 2, 3)])
def test_batch_sharding_over_contexts(double_exe, workdir, gpus, batch, threads):
    """-gpus N: every pipeline batch is cut into contiguous record-range shards with both mates of a pair on one context
    (SURVEY §8e); the output must not depend on N, the batch size or -t."""
    meta, fq, dump, expected = T.load_golden("paired_k31_maxcor8", workdir + "_cpu")
    od = os.path.join(workdir + "_cpu", "shard_%d_%d" % (gpus, batch))
    r = subprocess.run([double_exe, "-p"] + fq + ["-c", dump, "-k", str(meta["k"]), "-od", od, "-t", str(threads), "-gpus", str(gpus),
                                                  "-batch", str(batch)] + meta["extra"], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert T.parse_ref_stderr(r.stderr.decode()) == meta["stderr"]
    for f, exp in zip(meta["outputs"], expected):
        assert open(os.path.join(od, f[:-3]), "rb").read() == exp


def test_argv_surface(double_exe, workdir):
    """main.cpp:165-268: no arguments -> help, exit 0; unknown flag -> 'Unknown argument', exit 0; missing dump -> exit 1;
    options and files in any order."""
    r = subprocess.run([double_exe], capture_output=True)
    assert r.returncode == 0 and b"-k kmer_length" in r.stdout + r.stderr
    r = subprocess.run([double_exe, "-bogus"], capture_output=True)
    assert r.returncode == 0 and b"Unknown argument: -bogus" in r.stderr
    meta, fq, dump, expected = T.load_golden("single_repeat_n", workdir + "_cpu")
    r = subprocess.run([double_exe, "-r", fq[0], "-c", os.path.join(workdir, "no_such.dump"), "-k", str(meta["k"])], capture_output=True)
    assert r.returncode == 1 and b"Could not open file" in r.stderr
    od = os.path.join(workdir + "_cpu", "argv_order")
    args = ["-t", "2", "-od", od] + meta["extra"] + ["-k", str(meta["k"]), "-c", dump, "-r", fq[0]]  # files last, options first
    r = subprocess.run([double_exe] + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert open(os.path.join(od, meta["outputs"][0][:-3]), "rb").read() == expected[0]


@pytest.mark.skipif(not os.path.exists(T.REF_BIN), reason="oracle/_ref/rcorrector not built")
@pytest.mark.parametrize("gz", [False, True])
def test_fasta_input_matches_the_reference(double_exe, workdir, gz):
    """FASTA in -> <name>.cor.fa out (two lines per record, tags on the id line, no quality).  The reference's
    own FASTA path only works with -t >= 2 (its -t 1 path dereferences a NULL quality pointer): compare there."""
    import gzip

    meta, fq, dump, expected = T.load_golden("single_repeat_n", workdir + "_cpu")
    lines = open(fq[0], "rb").read().split(b"\n")
    fa = []
    for i in range(0, len(lines) - 1, 4):
        fa += [b">" + lines[i][1:], lines[i + 1]]
    w = os.path.join(workdir + "_cpu", "fasta_gz" if gz else "fasta")
    os.makedirs(w, exist_ok=True)
    path = os.path.join(w, "x.fa.gz" if gz else "x.fa")
    with (gzip.open if gz else open)(path, "wb") as f:
        f.write(b"\n".join(fa) + b"\n")
    res = []
    for tag, exe in (("ref", T.REF_BIN), ("ours", double_exe)):
        od = os.path.join(w, "o_" + tag)
        r = subprocess.run([exe, "-r", path, "-c", dump, "-k", str(meta["k"]), "-od", od, "-t", "2"] + meta["extra"], capture_output=True)
        assert r.returncode == 0, r.stderr
        out = os.path.join(od, "x.cor.fa.gz" if gz else "x.cor.fa")
        res.append(((gzip.open if gz else open)(out, "rb").read(), r.stderr))
    assert res[0][0] == res[1][0] and len(res[0][0]) > 0
    assert res[0][1] == res[1][1]  # Stored / rate / "Bad quality threshold is '\0'" / Processed lines, byte for byte


# ------------------------------------------------------------------------------------------------------------------
# The process surface against the reference binary itself: same return code, same stderr bytes, same
# stdout bytes, same output file names and (decompressed) contents — on the scenarios a user can reach from the command
# line.  Not reproduced, by decision: inputs on which the reference's output depends on stale buffer
# contents (blank lines inside a FASTQ file, a truncated last record), -t 0 (the reference then processes no reads), -k
# above 32 (the reference runs on with an overflowing mask; we refuse like run_rcorrector.pl does), -verbose traces.
def _surface_files(w):
    import gzip
    import shutil

    m1, fq1, dump1, _ = T.load_golden("paired_k23", w)
    m2, fq2, dump2, _ = T.load_golden("single_repeat_n", w)
    m3, fq3, dump3, _ = T.load_golden("paired_maxcor1", w)

    def cp(src, name):
        d = os.path.join(w, name)
        if not os.path.exists(d):
            shutil.copy(src, d)
        return d

    f = dict(a1=cp(fq1[0], "a_1.fq"), a2=cp(fq1[1], "a_2.fq"), b1=cp(fq3[0], "b_1.fq"), b2=cp(fq3[1], "b_2.fq"), s=cp(fq2[0], "s.fq"),
             dump=dump1, dump2=dump2, nope=os.path.join(w, "nope.fq"), nodump=os.path.join(w, "nope.dump"))
    lines = open(f["a2"], "rb").read().split(b"\n")
    f["short2"] = os.path.join(w, "short_2.fq")
    open(f["short2"], "wb").write(b"\n".join(lines[:400]) + b"\n")
    f["empty"] = os.path.join(w, "empty.fq")
    open(f["empty"], "wb").close()
    f["a2gz"] = os.path.join(w, "a_2.fq.gz")
    open(f["a2gz"], "wb").write(gzip.compress(open(f["a2"], "rb").read()))
    d = open(dump1, "rb").read().split(b"\n")
    f["odd_dump"] = os.path.join(w, "odd.dump")  # a count-1 record, a k-mer with N, a duplicate key with another count
    open(f["odd_dump"], "wb").write(b"\n".join(d[:200] + [b">1", d[1], b">5", b"ACGTNACGTACGTACGTACGTACG", b">7", d[1]] + d[200:]))
    sl = open(f["s"], "rb").read().split(b"\n")
    fa = []
    for i in range(0, len(sl) - 1, 4):
        fa += [b">" + sl[i][1:], sl[i + 1]]
    f["fa"] = os.path.join(w, "x.fa")
    open(f["fa"], "wb").write(b"\n".join(fa) + b"\n")
    return f


SURFACE = {
    "two_pairs": "-p {a1} {a2} -p {b1} {b2} -c {dump} -k 23 -od @OD -t 2",
    "pair_then_single": "-p {a1} {a2} -r {s} -c {dump} -k 23 -od @OD -t 2",
    "single_then_pair": "-r {s} -p {a1} {a2} -c {dump} -k 23 -od @OD -t 2",
    "two_singles_stdout": "-r {s} -r {a1} -c {dump} -k 23 -od @OD -t 2 -stdout",
    "stdout_pair": "-p {a1} {a2} -c {dump} -k 23 -od @OD -stdout -t 2",
    "options_after_files": "-r {s} -k 23 -c {dump} -od @OD",
    "same_file_twice": "-r {s} -r {s} -c {dump} -k 23 -od @OD",
    "pair_plain_and_gz": "-p {a1} {a2gz} -c {dump} -k 23 -od @OD -t 2",
    "unpaired_short_mate2": "-p {a1} {short2} -c {dump} -k 23 -od @OD -t 2",
    "unpaired_short_mate1": "-p {short2} {a2} -c {dump} -k 23 -od @OD -t 2",
    "empty_file": "-r {empty} -c {dump} -k 23 -od @OD",
    "missing_file": "-r {nope} -c {dump} -k 23 -od @OD",
    "missing_dump": "-r {s} -c {nodump} -k 23 -od @OD",
    "not_a_read_file": "-r {dump} -c {dump} -k 23 -od @OD",
    "od_nested_missing": "-r {s} -c {dump} -k 23 -od @OD/x/y",
    "no_args": "",
    "help": "-h",
    "unknown_flag": "-r {s} -zzz -c {dump}",
    "wk_half": "-r {a1} -c {dump} -k 23 -od @OD -wk 0.5",
    "wk_zero": "-r {a1} -c {dump} -k 23 -od @OD -wk 0",
    "wk_one": "-r {a1} -c {dump} -k 23 -od @OD -wk 1",
    "maxcor_is_dead": "-r {a1} -c {dump} -k 23 -od @OD -maxcor 1",
    "maxcorK_0": "-r {a1} -c {dump} -k 23 -od @OD -maxcorK 0",
    "maxcorK_20": "-r {a1} -c {dump} -k 23 -od @OD -maxcorK 20",
    "many_threads": "-p {a1} {a2} -c {dump} -k 23 -od @OD -t 64",
    "k_smaller_than_dump": "-p {a1} {a2} -c {dump} -k 21 -od @OD",
    "k_5": "-r {a1} -c {dump} -k 5 -od @OD",
    "dump_oddities": "-p {a1} {a2} -c {odd_dump} -k 23 -od @OD",
    "empty_dump": "-r {a1} -c {empty} -k 23 -od @OD",
    "fasta_interleaved": "-i {fa} -c {dump2} -k 23 -od @OD -t 2",
}


@pytest.mark.skipif(not os.path.exists(T.REF_BIN), reason="oracle/_ref/rcorrector not built")
@pytest.mark.parametrize("case", sorted(SURFACE))
def test_process_surface_matches_the_reference(double_exe, workdir, case):
    import gzip

    w = os.path.join(workdir + "_cpu", "surface")
    os.makedirs(w, exist_ok=True)
    f = _surface_files(w)
    res = []
    for tag, exe in (("ref", T.REF_BIN), ("ours", double_exe)):
        od = os.path.join(w, "%s_%s" % (case, tag))
        args = [a.replace("@OD", od) for a in SURFACE[case].format(**f).split()]
        r = subprocess.run([exe] + args, capture_output=True, cwd=w)
        files = {}
        if os.path.isdir(od):
            for name in sorted(os.listdir(od)):
                p = os.path.join(od, name)
                if os.path.isfile(p):
                    data = open(p, "rb").read()
                    files[name] = gzip.decompress(data) if name.endswith(".gz") and data else data
        res.append((r.returncode, r.stderr.replace(od.encode(), b"@OD"), r.stdout, files))
    assert res[0][0] == res[1][0], (res[0][1], res[1][1])
    assert res[0][1] == res[1][1]
    assert res[0][2] == res[1][2]
    if res[0][0] == 0:
        assert res[0][3] == res[1][3]
    else:
        # error exits: same files created; what they hold when the error strikes depends on the batch size (the reference has
        # written its 512 x t-record batches before an unpaired record, we the complete pairs before it — deterministically)
        assert sorted(res[0][3]) == sorted(res[1][3])
        if case.startswith("unpaired"):
            assert all(v.count(b"\n") == 400 for v in res[1][3].values())  # the 100 complete pairs, both files


def _gzip_members(data):
    """Split a gzip file into its members (RFC 1952 allows a file to be a concatenation of members)."""
    import zlib

    out = []
    while data:
        d = zlib.decompressobj(16 + zlib.MAX_WBITS)
        out.append(d.decompress(data))
        assert d.eof
        data = d.unused_data
    return out


def test_gz_output_is_written_as_parallel_members(double_exe, workdir):
    """*.gz in -> *.gz out: the pieces of a batch are compressed by the formatting threads as gzip members and written back
    to back.  The decompressed bytes equal the plain run's; the reference binary reads such a multi-member file as its own
    input (gzgets), so do gzip -d and Python; an input without records still leaves a valid (empty) gzip file."""
    import gzip

    meta, fq, dump, expected = T.load_golden("paired_k23", workdir + "_cpu")
    w = os.path.join(workdir + "_cpu", "gzmembers")
    os.makedirs(w, exist_ok=True)
    big = open(fq[0], "rb").read() * 24  # 12 000 records: more than one piece per batch with -t 4
    plain = os.path.join(w, "big.fq")
    open(plain, "wb").write(big)
    gz = os.path.join(w, "big.fq.gz")
    cut = big.find(b"\n@", len(big) // 2) + 1
    open(gz, "wb").write(gzip.compress(big[:cut], 1) + gzip.compress(big[cut:], 1))  # a two-member input as well
    outs = {}
    for tag, path, t in (("plain", plain, "4"), ("gz", gz, "4"), ("gz_t1", gz, "1")):
        od = os.path.join(w, "o_" + tag)
        r = subprocess.run([double_exe, "-r", path, "-c", dump, "-k", str(meta["k"]), "-od", od, "-t", t], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[tag] = open(os.path.join(od, "big.cor.fq" + (".gz" if tag != "plain" else "")), "rb").read()
    members = _gzip_members(outs["gz"])
    assert len(members) > 1 and b"".join(members) == outs["plain"]
    assert gzip.decompress(outs["gz_t1"]) == outs["plain"]
    if os.path.exists(T.REF_BIN):  # the reference reads our multi-member output with its own reader, and writes the same bytes
        od = os.path.join(w, "o_ref")
        r = subprocess.run([T.REF_BIN, "-r", gz, "-c", dump, "-k", str(meta["k"]), "-od", od, "-t", "2"], capture_output=True)
        assert r.returncode == 0
        assert gzip.decompress(open(os.path.join(od, "big.cor.fq.gz"), "rb").read()) == outs["plain"]
        again = os.path.join(w, "again.fq.gz")
        open(again, "wb").write(outs["gz"])
        od2 = os.path.join(w, "o_ref_again")
        r = subprocess.run([T.REF_BIN, "-r", again, "-c", dump, "-k", str(meta["k"]), "-od", od2, "-t", "2"], capture_output=True)
        assert r.returncode == 0 and b"Processed 12000 reads" in r.stderr
    # header-only FASTQ-like file with no complete record cannot be sniffed; an input of zero records after the first line:
    empty_gz = os.path.join(w, "none.fq.gz")
    open(empty_gz, "wb").write(gzip.compress(b"@only_an_id_line\n", 1))
    od = os.path.join(w, "o_none")
    r = subprocess.run([double_exe, "-r", empty_gz, "-c", dump, "-k", str(meta["k"]), "-od", od], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    data = open(os.path.join(od, "none.cor.fq.gz"), "rb").read()
    assert len(data) > 0 and isinstance(gzip.decompress(data), bytes)


def test_binary_stops_on_reads_the_gpu_path_does_not_reproduce(double_exe, workdir):
    """rcx_stats.unsupported > 0 after a batch: the binary exits 1 with a message instead of writing output that may differ."""
    meta, fq, dump, expected = T.load_golden("single_repeat_n", workdir + "_cpu")
    od = os.path.join(workdir + "_cpu", "unsupported")
    r = subprocess.run([double_exe, "-r", fq[0], "-c", dump, "-k", str(meta["k"]), "-od", od] + meta["extra"], capture_output=True,
                       env=dict(os.environ, RCX_DOUBLE_UNSUPPORTED="3"))
    assert r.returncode == 1 and b"3 reads of this batch need a case the GPU path does not reproduce" in r.stderr
    assert not os.path.exists(os.path.join(od, meta["outputs"][0][:-3]))  # no truncated *.cor.fq is left behind
    assert not os.path.exists(od)                                          # nor the directory this run created for it


@pytest.mark.parametrize("batch", [2, 4, 8, 16])
def test_tiny_batches_with_long_reads(double_exe, workdir, batch):
    """-batch below 16 with reads of up to 1 022 bases (found by tools/fuzz_binary_vs_ref.py: the batch buffers were sized from
    the record budget alone)."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir + "_cpu")
    w = os.path.join(workdir + "_cpu", "tiny_%d" % batch)
    os.makedirs(w, exist_ok=True)
    lines = open(fq[0], "rb").read().split(b"\n")
    long_seq = (lines[1] * 12)[:1022]
    lines[5], lines[7] = long_seq, b"I" * len(long_seq)
    src = os.path.join(w, "x.fq")
    open(src, "wb").write(b"\n".join(lines[:4 * 40]) + b"\n")
    outs = []
    for tag, b in (("tiny", str(batch)), ("big", "65536")):
        od = os.path.join(w, "o_" + tag)
        r = subprocess.run([double_exe, "-r", src, "-c", dump, "-k", str(meta["k"]), "-od", od, "-batch", b, "-t", "2"], capture_output=True)
        assert r.returncode == 0 and b"Processed 40 reads" in r.stderr, r.stderr.decode()
        outs.append(open(os.path.join(od, "x.cor.fq"), "rb").read())
    assert outs[0] == outs[1] and outs[0].count(b"\n") == 160


@pytest.mark.parametrize("threads", ["1", "4"])
def test_interleaved_file_without_final_newline_keeps_its_last_pair(double_exe, workdir, threads):
    """Found by tools/fuzz_binary_vs_ref.py (seed 9, case 191; kept as tests/golden/host_interleaved_no_final_newline): the
    mapped-file reader took the complete records (an odd number) and left the last, newline-less record to the serial reader,
    so the mates of the last pair were corrected without each other."""
    meta, fq, dump, expected = T.load_golden("paired_k23", workdir + "_cpu")
    fx = os.path.join(T.GOLDEN, "host_interleaved_no_final_newline")
    exp = open(os.path.join(fx, "expected.cor.fq"), "rb").read()
    w = os.path.join(workdir + "_cpu", "inter_nonl_" + threads)
    os.makedirs(w, exist_ok=True)
    for tag, env, batch in (("fast", {}, "1048576"), ("fast_tiny", {}, "8"), ("serial", {"RCX_SERIAL_IO": "1"}, "1048576")):
        od = os.path.join(w, tag)
        r = subprocess.run([double_exe, "-i", os.path.join(fx, "input.fq"), "-c", dump, "-k", str(meta["k"]), "-od", od, "-t", threads,
                            "-batch", batch], capture_output=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr.decode()
        assert b"Processed 16 reads" in r.stderr and b"Corrected 32 bases" in r.stderr
        assert open(os.path.join(od, "input.cor.fq"), "rb").read() == exp
