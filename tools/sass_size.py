#!/usr/bin/env python
"""Static SASS size of the correction kernel by source region (nvdisasm -g line info).
usage: python tools/sass_size.py  (needs rcorrector_b200/librcx.so built with -lineinfo)"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rcorrector_b200", "librcx.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, "rcx_correct.sm_100a.cubin")], stdout=subprocess.PIPE, text=True).stdout
cor = open(os.path.join(root, "rcorrector_b200/csrc/rcx_correct.cu")).read().split("\n")
# region markers: function starts / phase comments in rcx_correct.cu
marks = []
for i, l in enumerate(cor, 1):
    m = re.match(r"^(?:static )?__(?:device|global)__ .*?(\w+)\(", l)
    if m: marks.append((i, "fn:" + m.group(1)))
    m = re.match(r"^\s*PH\((\d)\);\s*//\s*(.*)", l)
    if m: marks.append((i, "  kernel ph%s %s" % (m.group(1), m.group(2))))
marks.sort()
def region(f, ln):
    if not f.endswith("rcx_correct.cu"): return "hdr:" + os.path.basename(f)
    r = "top"
    for a, n in marks:
        if a <= ln: r = n
        else: break
    return r
cur = ("?", 0); agg = collections.Counter(); byline = collections.Counter()
kern = sys.argv[2] if len(sys.argv) > 2 else ""
insec = True
for l in txt.split("\n"):
    ms = re.match(r"^\s*\.section\s+\.text\.(\S+?),", l)
    if ms: insec = kern in ms.group(1)
    if not insec: continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        agg[region(*cur)] += 1; byline[cur] += 1
tot = sum(agg.values())
print("total %d instr = %.1f KB" % (tot, tot * 16 / 1024))
for k, v in sorted(agg.items(), key=lambda x: -x[1]): print("%6d %5.1f KB  %s" % (v, v * 16 / 1024, k))
if len(sys.argv) > 1:
    for (f, ln), v in byline.most_common(int(sys.argv[1])):
        src = cor[ln - 1].strip()[:90] if f.endswith("rcx_correct.cu") else ""
        print("%5d %s:%d %s" % (v, os.path.basename(f), ln, src))
