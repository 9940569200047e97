"""CPU checks of bench.py's host side: the workload generator's shard semantics (one transcriptome, one table, a
different shard of reads per rank) and the reference arm's JSON line on a tiny workload, single-end and paired.
No GPU, no compute through librcx."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import rcx_testlib as T

ROOT = T.ROOT


def _synth(wd, name, *extra):
    if not os.path.exists(T.SYNTH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools")])
    p = os.path.join(wd, name)
    subprocess.check_call([T.SYNTH, "-seed", "1", "-ntx", "40", "-reads", "800", "-len", "100", "-k", "23", "-bin", p, "-table", p + "_t"] + list(extra),
                          stderr=subprocess.DEVNULL)
    return {x: np.fromfile(p + ext, dtype=dt) for x, ext, dt in
            (("seq", ".seq", np.uint8), ("off", ".off", np.uint32), ("codes", "_t.codes", np.uint64), ("counts", "_t.counts", np.int32))}


def test_read_shards_share_the_table(workdir):
    a = _synth(workdir, "s_a")
    b = _synth(workdir, "s_b", "-rseed", "2")
    c = _synth(workdir, "s_c", "-rseed", "1")
    assert np.array_equal(a["codes"], b["codes"]) and np.array_equal(a["counts"], b["counts"])  # same transcriptome, same table
    assert len(a["off"]) == len(b["off"])
    assert not np.array_equal(a["seq"][:2000], b["seq"][:2000])  # another shard of reads
    assert np.array_equal(a["seq"], c["seq"]) and np.array_equal(a["off"], c["off"])  # -rseed defaults to -seed


def test_paired_shard_is_interleaved(workdir):
    p = _synth(workdir, "s_p", "-paired")
    assert len(p["off"]) - 1 == 1600  # 800 pairs -> 1600 records, mates at 2i, 2i+1
    lens = np.diff(p["off"].astype(np.int64))
    assert (lens[1::2] == 100).all()  # only mate 1 is ever truncated


@pytest.mark.parametrize("paired", [False, True])
def test_reference_arm_line(workdir, paired):
    if not os.path.exists(T.REF_BIN):
        pytest.skip("oracle/_ref/rcorrector not built")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "20000", "--tx-bases", "150000",
           "--cpu-sample", "1000", "--steps", "2", "--warmup", "1", "--workdir", os.path.join(workdir, "bw")]
    if paired:
        cmd += ["--paired", "--len", "150"]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["unit"] == "reads/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["paired"] is paired


def test_stdout_carries_only_the_json_line(workdir):
    """bench.py points file descriptor 1 at stderr until its one JSON line is due (NCCL's version banner, child processes)."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; g = bench.StdoutToStderr(); os.system('echo from-a-child'); "
            "print('from-python'); g.restore(); print('{\"ok\": 1}')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    assert r.stdout == '{"ok": 1}\n'
    assert "from-a-child" in r.stderr and "from-python" in r.stderr


@pytest.mark.parametrize("paired", [False, True])
def test_bench_prepass_scalars_and_sample_match_the_reference(workdir, paired):
    """bench.py derives ERROR_RATE's input order and badQualityThreshold itself (no hard-coded parameters): the records it
    hands to rcx_estimate_error_rate are the text dump's leading records, its quality threshold is the one the reference
    binary prints for the sample file, and the FASTQ sample is the prefix of read seed 1's shard."""
    if not os.path.exists(T.REF_BIN):
        pytest.skip("oracle/_ref/rcorrector not built")
    sys.path.insert(0, ROOT)
    import bench

    argv = ["--reads", "6000", "--tx-bases", "60000", "--cpu-sample", "1000", "--workdir", os.path.join(workdir, "bw2")]
    if paired:
        argv += ["--paired", "--len", "150"]
    a = bench.parse_args(argv)
    os.makedirs(a.workdir, exist_ok=True)
    td = bench.gen_table(a, want_dump=True, fq_limit=a.cpu_sample)
    codes = np.fromfile(os.path.join(td, "table.codes"), dtype=np.uint64)
    counts = np.fromfile(os.path.join(td, "table.counts"), dtype=np.int32)
    dcodes, dcounts = T.dump_to_arrays(os.path.join(td, "table.dump"), a.k)
    pre = bench.dump_order_prefix(codes, counts)
    assert np.array_equal(pre, dcodes[: len(pre)]) and len(pre) == len(dcodes)  # (few records >= 1000 here: the whole order)
    (rs, nr), = bench.shard_plan(a, 1, 0)
    sd = bench.gen_shard(a, rs, nr)
    qual = np.fromfile(os.path.join(sd, "reads.qual"), dtype=np.uint8)
    seq = np.fromfile(os.path.join(sd, "reads.seq"), dtype=np.uint8)
    off = np.fromfile(os.path.join(sd, "reads.off"), dtype=np.uint32)
    bs = [T.read_fastq(os.path.join(td, f)) for f in (("sample.fq", "sample2.fq") if paired else ("sample.fq",))]
    b = T.interleave(*bs) if paired else bs[0]
    assert b.n == a.cpu_sample and np.array_equal(b.seq, seq[: off[b.n]]) and np.array_equal(b.off, off[: b.n + 1])
    r = bench.run_reference_binary(a, td, bench.reference_sample_files(a, td, 1), 2, keep_first=True)
    assert bench.bad_quality_threshold(qual, off, paired, a.cpu_sample) == r["bad_qual"]
    meta = json.load(open(os.path.join(bench.keep_dir(td), "meta.json")))
    assert meta["reads"] == a.cpu_sample and meta["bad_qual"] == r["bad_qual"]
    # the kept *.cor.fq is what the oracle gives for the sample with the printed parameters
    table = T.OracleTable(a.k)
    table.put(codes, counts)
    er = table.error_rate(os.path.join(td, "table.dump"))
    assert "%.6f" % er == r["rate6"]
    so, corr, l, m, h = T.oracle_correct(table, b, a.k, T.INTERLEAVED if paired else T.SINGLE, er, r["bad_qual"], a.maxcork)
    for i in range(2 if paired else 1):
        want = open(os.path.join(bench.keep_dir(td), "sample%s.cor.fq" % ("" if i == 0 else "2")), "rb").read()
        assert T.format_fastq(b, so, corr, l, m, h, rows=range(i, b.n, 2) if paired else None) == want
