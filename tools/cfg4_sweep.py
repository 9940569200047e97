#!/usr/bin/env python
"""BASELINE.json configs[3]: table-size sweep of the lookup kernel (K1) and the kernels behind it, up to 1 B distinct
k-mers, k = 23, 150-bp reads.   usage (GPU box): python tools/cfg4_sweep.py [--kmers 64,256,1000] [--reads 20000000]

Everything is generated ON THE GPU (no 27 GB text dump, no host arrays): a random "transcriptome" of G bases whose G-k+1
windows are the table's k-mers (distinct for k = 23 up to a handful of collisions: 4^23 = 7e13), counts following a
log-normal expression level per ~1 500-base stretch; reads are 150-bp windows of the same sequence with 1 % substitutions, so the scan meets stored contexts on the
clean windows and absent ones around every error, like the bench workload does.  The reads go through the WHOLE device
path (rcx_correct_batch_device), so K1 runs in the mode it runs in production (thresholds, triage, preparation fused);
its CUDA-event time, its sector requests and the random-request rate of the same table (rcx_measure_random_read) are
reported per table size, with the search and finalize kernels beside it.

A table of unrelated random k-mers would be the wrong model: every k-mer would bring two private (k-1)-mer contexts and
the context-keyed table would be twice as large as for any sequence-derived k-mer set.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmers", default="64,256,1000", help="table sizes in millions of k-mers")
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--k", type=int, default=23)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch

    from rcorrector_b200 import SINGLE, Corrector

    dev = torch.device("cuda", a.device)
    torch.cuda.set_device(dev)
    k, L, n = a.k, a.len, a.reads
    res = {"config": "configs[3]-shaped: table-size sweep, k=%d, %d x %d bp reads, %.3g substitutions per base" % (k, n, L, a.err),
           "gpu": torch.cuda.get_device_name(dev), "device": a.device, "sizes": []}
    for mk in [int(x) for x in a.kmers.split(",")]:
        nk = mk * 1_000_000
        G = nk + k - 1
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + mk)
        t0 = time.time()
        genome = torch.randint(0, 4, (G,), dtype=torch.uint8, device=dev, generator=g)
        codes = torch.zeros(nk, dtype=torch.int64, device=dev)
        for j in range(k):  # forward code of window i: first base in the highest bits (KmerCode::Append)
            codes.mul_(4).add_(genome[j:j + nk])
        # counts like a transcriptome's: one expression level per ~1 500-base "transcript" (log-normal), +-15 % noise per k-mer,
        # never below 2 -- neighbouring windows of a clean read agree, so clean reads leave at the triage like real ones do
        nblk = (nk + 1499) // 1500
        level = torch.exp(3.3 + 1.3 * torch.randn(nblk, device=dev, generator=g))
        counts = level.repeat_interleave(1500)[:nk] * (0.85 + 0.3 * torch.rand(nk, device=dev, generator=g))
        counts = counts.clamp_(2, 2e9).round_().to(torch.int32)
        del level
        # reads: windows of the sequence, substitutions at rate err, constant quality
        starts = torch.randint(0, G - L, (n,), dtype=torch.int64, device=dev, generator=g)
        seq = torch.empty((n, L), dtype=torch.uint8, device=dev)
        ar = torch.arange(L, device=dev)
        step = 2_000_000
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            b = genome[starts[lo:hi, None] + ar[None, :]]
            e = torch.rand((hi - lo, L), device=dev, generator=g) < a.err
            b = torch.where(e, (b + torch.randint(1, 4, b.shape, dtype=torch.uint8, device=dev, generator=g)) & 3, b)
            seq[lo:hi] = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[b.long()]
        del starts, genome
        seq0 = seq.reshape(-1)
        qual = torch.full_like(seq0, ord("I"))
        off = (torch.arange(n + 1, dtype=torch.int64, device=dev) * L).view(torch.int32)[::2].contiguous()  # low words = the uint32 offsets
        torch.cuda.synchronize()
        t_gen = time.time() - t0
        cor = Corrector(a.device)
        t0 = time.time()
        cor.build_table(k, codes, counts, n=nk, on_device=True)
        cor.sync()
        t_build = time.time() - t0
        del codes, counts
        torch.cuda.empty_cache()
        info = cor.table_info()
        cor.set_params(k, 0.01, ord("5"), 4)
        rr = cor.measure_random_read(32)
        work = torch.empty_like(seq0)
        outs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
        stream = torch.cuda.ExternalStream(cor.stream(), device=dev)
        sts = []
        for it in range(a.warmup + a.steps):
            with torch.cuda.stream(stream):
                work.copy_(seq0, non_blocking=True)
            cor.correct_device(SINGLE, n, work, qual, off, n * L, L, *outs)
            st = cor.stats()
            if it >= a.warmup:
                sts.append(st)
        ms = lambda key: sum(s[key] for s in sts) / len(sts)
        st = sts[-1]
        req_peak = rr / 32.0
        row = {"kmers": info["entries"], "table_bytes": info["bytes"], "bytes_per_kmer": info["bytes"] / info["entries"],
               "sectors": info["slots"], "gen_s": round(t_gen, 1), "build_s": round(t_build, 2),
               "random_request_peak_grps": req_peak, "unsupported": st["unsupported"], "work_items": st["work_items"],
               "corrected_reads": int((outs[0] > 0).sum().item()), "kernels": {}}
        for name, tkey, rkey in (("rcx_scan_thread_kernel", "ms_scan", "scan_probes"), ("rcx_dfs_kernel", "ms_dfs", "search_probes"),
                                 ("rcx_finalize_thread_kernel", "ms_finalize", "rescan_probes")):
            t = ms(tkey)
            row["kernels"][name] = {"ms": t, "requests": st[rkey], "grequests_per_s": st[rkey] / t / 1e6,
                                    "frac_of_random_request_peak": st[rkey] / t / 1e6 / req_peak,
                                    "dram_line_model_gbs": 128 * st[rkey] / t / 1e6}
        row["reads_per_s"] = n / (ms("ms_scan") + ms("ms_correct")) * 1e3
        res["sizes"].append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
        cor.close()
        del seq, seq0, qual, off, work, outs
        torch.cuda.empty_cache()
    txt = json.dumps(res, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
