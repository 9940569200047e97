"""CPU: the C-ABI library loads and exports every symbol include/rcx.h declares; without a GPU it fails
loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rcx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rcx_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from rcorrector_b200 import api

    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "librcx.so does not export %s" % s
    assert sorted(api.SYMBOLS) == syms
    assert lib.rcx_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    from rcorrector_b200 import Corrector, RcxError

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the failure path is not reachable")
    with pytest.raises(RcxError, match="no CUDA device"):
        Corrector(0)


def test_product_does_not_link_or_import_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "rcorrector_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for token in ("rcx_oracle", "librcx_oracle", "oracle/", "orc_"):
                    assert token not in txt, "%s references the oracle (%s)" % (f, token)
    out = os.popen("ldd %s 2>/dev/null" % os.path.join(pkg, "librcx.so")).read()
    assert "oracle" not in out


def test_host_binary_cli_surface():
    """main.cpp:165-247: no arguments -> help on stderr, exit 0; unknown flag -> message, exit 0."""
    import subprocess

    exe = os.path.join(ROOT, "rcorrector_b200", "bin", "rcorrector")
    if not os.path.exists(exe):
        pytest.skip("bin/rcorrector not built")
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 0 and b"Usage: ./rcorrector [OPTIONS]" in r.stderr
    r = subprocess.run([exe, "-bogus"], capture_output=True)
    assert r.returncode == 0 and b"Unknown argument: -bogus" in r.stderr
    r = subprocess.run([exe, "-c", "/nonexistent/dump", "-r", "x.fq"], capture_output=True)
    assert r.returncode == 1 and b"Could not open file /nonexistent/dump" in r.stderr


def test_product_binary_fails_loudly_without_gpu(workdir):
    """The shipped binary links librcx.so only (never the oracle or the test double): on a machine without a CUDA device a
    valid command line ends with exit code 1 and the library's message, not with output files from some CPU path."""
    import subprocess

    import torch

    import rcx_testlib as T

    exe = os.path.join(ROOT, "rcorrector_b200", "bin", "rcorrector")
    if not os.path.exists(exe):
        pytest.skip("bin/rcorrector not built")
    out = os.popen("ldd %s 2>/dev/null" % exe).read()
    assert "librcx.so" in out and "oracle" not in out
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the failure path is not reachable")
    meta, fq, dump, expected = T.load_golden("single_repeat_n", workdir)
    od = os.path.join(workdir, "nogpu_out")
    r = subprocess.run([exe, "-r", fq[0], "-c", dump, "-k", str(meta["k"]), "-od", od], capture_output=True)
    assert r.returncode == 1 and b"no CUDA device" in r.stderr
    assert all(os.path.getsize(os.path.join(od, f)) == 0 for f in os.listdir(od))
