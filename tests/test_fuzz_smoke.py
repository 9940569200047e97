"""CPU: a short run of the differential campaigns kept under tools/ (the long runs are recorded in profiles/r01_z_fuzz_*):
the oracle against the reference binary on random workloads, and the binary's host side (librcx test double) against the
reference binary on random awkward FASTQ / FASTA input.  Only where oracle/_ref/rcorrector exists."""
import os
import random
import shutil
import sys

import pytest

import rcx_testlib as T

sys.path.insert(0, os.path.join(T.ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.exists(T.REF_BIN), reason="oracle/_ref/rcorrector not built")


def test_oracle_campaign_smoke(workdir):
    import fuzz_oracle_vs_ref as F

    rng = random.Random(20261017)
    for i in range(12):
        wd = os.path.join(workdir, "fz_oracle_%d" % i)
        os.makedirs(wd)
        problems, desc, stats, name = F.one_case(rng, wd, i)
        assert not problems, (problems, desc)
        shutil.rmtree(wd, ignore_errors=True)


def test_binary_campaign_smoke(workdir):
    import fuzz_binary_vs_ref as F

    top = os.path.join(workdir, "fz_binary")
    os.makedirs(top)
    exe = F.build_double(top)
    meta, fq, dump, _ = T.load_golden("paired_k23", top)
    base = F.load_pairs(fq)
    rng = random.Random(20261017)
    for i in range(30):
        wd = os.path.join(top, "c%03d" % i)
        os.makedirs(wd)
        same, desc, res = F.one_case(rng, exe, base, dump, meta["k"], wd)
        assert same, desc
        shutil.rmtree(wd, ignore_errors=True)
