#!/usr/bin/env python
"""End-to-end timing of the drop-in `rcorrector` binary against the reference binary on the same generated files
(GPU box).  Prints one JSON object: per-phase wall times from the timestamps of the stderr lines both binaries print
(`Stored`, `Weak kmer`, `Bad quality`, `Processed`), output md5s (must be equal).
usage: python tools/binary_round.py [reads] [ntx] [len] [paired 0/1] [extra args for ours...]
       several variants of ours in one run (the reference runs once): ... [paired 0/1] -- -gpus 1 -- -gpus 8"""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYNTH = os.path.join(ROOT, "tools", "bin", "rcx_synth")
REF = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
OURS = os.path.join(ROOT, "rcorrector_b200", "bin", "rcorrector")


def run_timed(cmd):
    t0 = time.time()
    p = subprocess.Popen(cmd, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL)
    marks, buf = {}, b""
    while True:
        ch = p.stderr.read(1)
        if not ch:
            break
        buf += ch
        if ch == b"\n":
            line = buf.decode(errors="replace")
            buf = b""
            now = time.time() - t0
            for key, tag in (("Stored", "stored"), ("Weak kmer", "rate"), ("Bad quality", "badq"), ("Processed", "processed")):
                if line.startswith(key):
                    marks[tag] = now
                    marks[tag + "_line"] = line.strip()
    p.wait()
    marks["exit"] = time.time() - t0
    marks["rc"] = p.returncode
    return marks


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    ntx = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
    rlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    paired = len(sys.argv) > 4 and sys.argv[4] == "1"
    extra = sys.argv[5:]
    variants = []
    if "--" in extra:  # several argument sets for ours
        cur = None
        for x in extra:
            if x == "--":
                cur = []
                variants.append(cur)
            elif cur is not None:
                cur.append(x)
    else:
        variants = [extra]
    d = "/dev/shm/rcx_binary_round"
    os.makedirs(d, exist_ok=True)
    k = 23
    cmd = [SYNTH, "-seed", "3", "-k", str(k), "-ntx", str(ntx), "-txmin", "300", "-txmax", "2500", "-polya", "0.3", "-sigma", "2.6", "-reads",
           str(reads), "-len", str(rlen), "-err", "0.01", "-nrate", "0.002", "-short", "0.02", "-fq1", os.path.join(d, "r_1.fq"), "-dump",
           os.path.join(d, "t.dump")]
    if paired:
        cmd += ["-paired", "-fq2", os.path.join(d, "r_2.fq")]
    t0 = time.time()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_s = time.time() - t0
    files = ["-p", os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")] if paired else ["-r", os.path.join(d, "r_1.fq")]
    cores = os.cpu_count() or 1
    res = {"reads": reads * (2 if paired else 1), "len": rlen, "k": k, "cores": cores, "generate_s": gen_s,
           "dump_bytes": os.path.getsize(os.path.join(d, "t.dump")), "fastq_bytes": os.path.getsize(os.path.join(d, "r_1.fq")) * (2 if paired else 1)}
    outs = {}
    runs = [("reference", REF, [])] + [("ours" if i == 0 else "ours_%d" % i, OURS, v) for i, v in enumerate(variants)]
    for name, binary, more in runs:
        od = os.path.join(d, "out_" + name)
        os.makedirs(od, exist_ok=True)
        m = run_timed([binary] + files + ["-c", os.path.join(d, "t.dump"), "-k", str(k), "-od", od, "-t", str(cores)] + more)
        m["load_table_s"] = m.get("stored", 0)
        m["error_rate_s"] = m.get("rate", 0) - m.get("stored", 0)
        m["bad_quality_s"] = m.get("badq", 0) - m.get("rate", 0)
        m["correct_s"] = m.get("processed", 0) - m.get("badq", 0)
        m["reads_per_s_correct_phase"] = res["reads"] / max(m["correct_s"], 1e-9)
        m["reads_per_s_whole_run"] = res["reads"] / max(m["exit"], 1e-9)
        outs[name] = sorted(os.path.join(od, f) for f in os.listdir(od))
        m["md5"] = [md5(f) for f in outs[name]]
        m["args"] = more
        res[name] = m
    res["identical_output"] = all(res["reference"]["md5"] == res[n]["md5"] for n, _, _ in runs[1:])
    res["speedup_correct_phase"] = res["reference"]["correct_s"] / max(res["ours"]["correct_s"], 1e-9)
    res["speedup_whole_run"] = res["reference"]["exit"] / max(res["ours"]["exit"], 1e-9)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
