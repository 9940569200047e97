#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (ncu --set full) for profiles/: duration, DRAM bytes, issue utilisation,
warp stall reasons per issued instruction.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("## %s\n" % name)
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        print("\nWarp stall reasons (cycles per issued instruction, > 0.2):\n")
        print("| reason | ratio |\n|---|---|")
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.2:
                    print("| %s | %.2f |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        print()


if __name__ == "__main__":
    main()
