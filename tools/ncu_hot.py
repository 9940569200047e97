#!/usr/bin/env python
"""Top source lines by executed warp instructions / stall samples for each kernel of an .ncu-rep captured with
--import-source on (-lineinfo build).  usage: python tools/ncu_hot.py rep.ncu-rep [N]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    fpath, func, hdr = None, None, None
    per = collections.OrderedDict()  # func -> {(file, line, src): [samples, instr]}
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1].split("/")[-1]
        elif row[0] == "Function Name":
            func = row[1]
        elif row[0] == "Line No":
            hdr = row
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif row[0].isdigit() and hdr:
            try:
                smp, ins = float(row[i_s] or 0), float(row[i_i] or 0)
            except ValueError:
                continue
            d = per.setdefault(func, collections.OrderedDict())
            key = (fpath, int(row[0]), row[1].strip())
            v = d.setdefault(key, [0.0, 0.0])
            v[0] += smp
            v[1] += ins
    for func, d in per.items():
        tot = sum(v[0] for v in d.values()) or 1
        toti = sum(v[1] for v in d.values()) or 1
        print("## %s  (samples %.0f, warp instructions %.3g)\n" % (func, tot, toti))
        print("| instr % | samples % | where | source |\n|---|---|---|---|")
        for (f, ln, src), v in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
            print("| %.1f | %.1f | %s:%d | `%s` |" % (100 * v[1] / toti, 100 * v[0] / tot, f, ln, src[:120].replace("|", "\\|")))
        print()


if __name__ == "__main__":
    main()
