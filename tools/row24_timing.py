#!/usr/bin/env python
"""The experimental scan kernel with 24-bit shared-memory count rows (RCX_ROW24=1, rcx_scan_thread24.cu) against the shipped one
(32-bit rows) on the configs[2]-shaped workload (paired 150 bp, the 197.6 M k-mer table, whose largest count escapes the
24-bit rows): kernel times from the library's CUDA events, and every output of the two runs compared
(profiles/r02_u_row24_experiment.md).
usage (GPU box): python tools/row24_timing.py [--subshards 1] [--fast] [--out file.json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subshards", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--fast", action="store_true", help="skip the pre-pass: error rate 0.01 and quality threshold 48 as bench.py derives them for this workload")
    ap.add_argument("--out", default=None)
    o = ap.parse_args()
    import torch

    from rcorrector_b200 import INTERLEAVED, Corrector

    a = bench.parse_args(["--config", "3"])
    dev = torch.device("cuda", 0)
    td = bench.gen_table(a)
    codes = np.fromfile(os.path.join(td, "table.codes"), dtype=np.uint64)
    counts = np.fromfile(os.path.join(td, "table.counts"), dtype=np.int32)
    cor = Corrector(0)
    cor.build_table(a.k, codes, counts)
    cor.sync()
    er = 0.01 if o.fast else cor.estimate_error_rate(bench.dump_order_prefix(codes, counts), 0.95)
    big = int((counts >= (1 << 21)).sum())
    res = {"workload": "configs[2] shape: %d sub-shard(s) of 6 M paired 150-bp reads, table of %d k-mers" % (o.subshards, len(codes)),
           "largest_count": int(counts.max()), "kmers_with_count_ge_2^21": big, "runs": {}}
    del codes, counts
    plan = bench.shard_plan(a, 1, 0)[: o.subshards]
    dirs = [bench.gen_shard(a, rs, nr) for rs, nr in plan]
    seqs = [np.fromfile(os.path.join(d, "reads.seq"), dtype=np.uint8) for d in dirs]
    quals = [np.fromfile(os.path.join(d, "reads.qual"), dtype=np.uint8) for d in dirs]
    offs = [np.fromfile(os.path.join(d, "reads.off"), dtype=np.uint32).astype(np.int64) for d in dirs]
    base, cat = 0, [np.zeros(1, dtype=np.int64)]
    for x in offs:
        cat.append(x[1:] + base)
        base += int(x[-1])
    off = np.concatenate(cat)
    n, total, max_len = len(off) - 1, int(off[-1]), int(np.diff(off).max())
    bq = 48 if o.fast else bench.bad_quality_threshold(quals[0], offs[0].astype(np.uint32), True, min(a.cpu_sample, n))
    cor.set_params(a.k, er, bq, a.maxcork)
    d_seq0 = torch.from_numpy(np.concatenate(seqs)).to(dev)
    d_qual = torch.from_numpy(np.concatenate(quals)).to(dev)
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    d_seq = torch.empty_like(d_seq0)
    outs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
    stream = torch.cuda.ExternalStream(cor.stream(), device=dev)
    ref = None
    for name, env in (("rows32", "0"), ("rows24", "1"), ("rows32 again", "0"), ("rows24 again", "1")):
        os.environ["RCX_ROW24"] = env
        sts = []
        for it in range(1 + o.steps):
            with torch.cuda.stream(stream):
                d_seq.copy_(d_seq0, non_blocking=True)
            cor.correct_device(INTERLEAVED, n, d_seq, d_qual, d_off, total, max_len, *outs)
            st = cor.stats()
            if it:
                sts.append(st)
        torch.cuda.synchronize()
        cur = [x.clone() for x in outs] + [d_seq.clone()]
        if ref is None:
            ref = cur
        differ = sum(int((x != y).sum().item()) for x, y in zip(cur, ref))
        ms = lambda key: sum(s[key] for s in sts) / len(sts)
        res["runs"][name] = {"ms_scan": round(ms("ms_scan"), 3), "ms_dfs": round(ms("ms_dfs"), 3), "ms_finalize": round(ms("ms_finalize"), 3),
                             "unsupported_flag": int(sts[-1]["unsupported"]), "outputs_that_differ_from_rows32": differ, "scan_requests": int(sts[-1]["scan_probes"]),
                             "reads": n}
        print(name, json.dumps(res["runs"][name]), file=sys.stderr, flush=True)
    txt = json.dumps(res, indent=1)
    if o.out:
        os.makedirs(os.path.dirname(os.path.abspath(o.out)), exist_ok=True)
        open(o.out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
