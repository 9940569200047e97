#!/usr/bin/env python
"""Per-sub-shard kernel times of bench.py's configs[2]-shaped job on ONE GPU: which of the 8 sub-shards (6 M reads each) costs
the search kernel more than the others, and why (work items, heavy items, search nodes).   usage: python tools/shard_variation.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from rcorrector_b200 import INTERLEAVED, Corrector

    only = [int(x) for x in sys.argv[1:] if x.isdigit()]  # read seeds to run (default: all eight)
    a = bench.parse_args(["--config", "3"])
    dev = torch.device("cuda", 0)
    td = bench.gen_table(a)
    codes = np.fromfile(os.path.join(td, "table.codes"), dtype=np.uint64)
    counts = np.fromfile(os.path.join(td, "table.counts"), dtype=np.int32)
    cor = Corrector(0)
    cor.build_table(a.k, codes, counts)
    cor.set_params(a.k, 0.01, ord("0"), a.maxcork)
    rows = []
    for rs, nr in bench.shard_plan(a, 1, 0):
        if only and rs not in only:
            continue
        d = bench.gen_shard(a, rs, nr)
        seq = torch.from_numpy(np.fromfile(os.path.join(d, "reads.seq"), dtype=np.uint8)).to(dev)
        qual = torch.from_numpy(np.fromfile(os.path.join(d, "reads.qual"), dtype=np.uint8)).to(dev)
        off_h = np.fromfile(os.path.join(d, "reads.off"), dtype=np.uint32)
        off = torch.from_numpy(off_h.view(np.int32)).to(dev)
        n = len(off_h) - 1
        res = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
        work = torch.empty_like(seq)
        stream = torch.cuda.ExternalStream(cor.stream(), device=dev)
        best = None
        for it in range(3):
            with torch.cuda.stream(stream):
                work.copy_(seq, non_blocking=True)
            cor.correct_device(INTERLEAVED, n, work, qual, off, int(off_h[-1]), 150, *res)
            st = cor.stats()
            if best is None or st["ms_dfs"] < best["ms_dfs"]:
                best = st
        row = {"read_seed": rs, "ms_scan": round(best["ms_scan"], 2), "ms_dfs": round(best["ms_dfs"], 2), "ms_finalize": round(best["ms_finalize"], 2),
               "work_items": best["work_items"], "heavy_items": best["heavy_items"], "search_nodes": best["search_nodes"],
               "search_probes": best["search_probes"], "second_pass": best["second_pass"]}
        if "diag" in os.environ.get("RCX_LIB", ""):  # -DRCX_DIAG build: walk length against untrusted-window count (buckets of 32)
            dc = [int(x) for x in cor.debug_counters(36)[17:30]]  # RCX_C_DIAG0 = 17
            row["diag"] = {"max_nodes_of_one_item": dc[0], "items_by_untrusted_windows_0_32_64_96": dc[1:5],
                           "items_over_2000_nodes": dc[5:9], "their_nodes": dc[9:13]}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del seq, qual, off, work, res
    cor.close()


if __name__ == "__main__":
    main()
