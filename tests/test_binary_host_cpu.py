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


def test_modes(workdir):
    G.test_binary_modes_match(workdir + "_cpu")


@pytest.mark.parametrize("case", ["no_final_newline", "ragged_quality", "crlf", "odd_interleaved", "plus_line_text", "regular_t1"])
def test_readers(case, workdir):
    G.test_parallel_reader_equals_serial_reader(case, workdir + "_cpu")


def test_dump_parsers(workdir):
    G.test_dump_parsers_agree(workdir + "_cpu")


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
