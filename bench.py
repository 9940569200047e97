#!/usr/bin/env python
"""bench.py — headline benchmark of the stage-3 correction hot path.

Workloads (BASELINE.json configs; `--config auto` picks by the GPU count):
  N = 1  configs[1]: 10 M synthetic single-end 100 bp reads, k = 23, ~50 M distinct k-mers, one B200.
  N > 1  configs[2]-shaped, STRONG scaling: ONE fixed set of 48 M paired-end 150 bp reads (24 M pairs, interleaved; 8
         sub-shards of 6 M reads) against a ~200 M-k-mer table; rank r of N corrects sub-shards [8r/N, 8(r+1)/N), the
         table is broadcast once over NCCL from rank 0 and inserted locally (timed, reported beside the line).  The
         same workload runs on one GPU with `--config 3`; the default N = 1 run appends it as `scaling_base_cfg3` so the
         1 -> N curve has its base on the same box.
  --config 5  configs[4]-shaped (scaled): k = 31, -maxcorK 8, 5 % errors, 150 bp.
A "step" is one pass of the hot path over the rank's reads: scan (+ thresholds, triage, first preparation) -> search ->
[second-pass preparation -> search] -> finalize.

  value : reads/s, whole job, inputs already resident in HBM (rcx_correct_batch_device)
  e2e   : reads/s through the C ABI with HOST buffers (rcx_correct_batch_compact: pinned host -> device copies of the
          bases + one quality bit per base, kernels, device -> host copies of correction/l/m/h + the applied edits, all
          inside the timed region; the library streams the batch through 3 pipeline slots so copies overlap kernels)
  roofline : dominant kernel (and the lookup kernel): algorithmic bytes (32 B per GetCount + bases) / CUDA-event time
          against MEASURED_PEAKS.json's HBM copy GB/s; table-sector requests/s against the random-request rate measured
          over the same table in the same run
  parity : the first --cpu-sample reads corrected through BOTH host interfaces, formatted as FASTQ and compared byte for
          byte with the *.cor.fq the UNMODIFIED reference binary wrote for them (kept by the reference arm of the same
          box, or produced by this run's cpu_baseline leg); the run aborts on a mismatch
  cpu_baseline : oracle/_ref/rcorrector -t <all host cores> on that sample against the full-size table, timed from its
          "Bad quality threshold" stderr line to "Processed ... reads" (SURVEY §8d); rank 0, N = 1 only

--impl reference: times the reference CPU binary itself on the same config (rank 0 only) and keeps its output for the
parity check of the other arm.
"""
import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SYNTH = os.path.join(ROOT, "tools", "bin", "rcx_synth")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "rcorrector")
METRIC = "reads/s (stage-3 correction, k=%d)"
NSUB = 8  # sub-shards of the fixed configs[2] read set


def log(*a):
    print(*a, file=sys.stderr, flush=True)


PRESETS = {
    # reads = records in the whole job (a pair = 2 records)
    "2": dict(reads=10_000_000, len=100, k=23, tx_bases=50_000_000, err=0.01, paired=False, maxcork=4, scaling="weak"),
    "3": dict(reads=48_000_000, len=150, k=23, tx_bases=200_000_000, err=0.01, paired=True, maxcork=4, scaling="strong"),
    "5": dict(reads=8_000_000, len=150, k=31, tx_bases=50_000_000, err=0.05, paired=False, maxcork=8, scaling="weak"),
}


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto", "2", "3", "5", "custom"],
                    help="BASELINE.json config: auto = 2 on one GPU, 3 (strong scaling) on several")
    ap.add_argument("--reads", type=int, default=None, help="records of the whole job (overrides the preset)")
    ap.add_argument("--len", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--tx-bases", type=int, default=None, help="transcriptome size ~= distinct k-mers")
    ap.add_argument("--err", type=float, default=None)
    ap.add_argument("--paired", action="store_true", default=None, help="mates interleaved (the -p / -i shape)")
    ap.add_argument("--maxcork", type=int, default=None, help="MAX_FIX_PER_K (-maxcorK)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="reads in the CPU reference / parity sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true", help="N = 1: skip the appended configs[2]-shaped measurement")
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args(argv)
    resolve_config(a)
    return a


def resolve_config(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = max(world, a.gpus)
    explicit = [x is not None for x in (a.reads, a.len, a.k, a.tx_bases, a.err, a.paired, a.maxcork)]
    if a.config == "auto":
        a.config = "custom" if any(explicit) else ("2" if n == 1 else "3")
    base = dict(PRESETS.get(a.config, PRESETS["2"]))
    for key in ("reads", "len", "k", "tx_bases", "err", "paired", "maxcork", "scaling"):
        v = getattr(a, key)
        setattr(a, key, base[key] if v is None else v)
    a.paired = bool(a.paired)
    if a.paired:
        a.reads &= ~1
    a.cpu_sample = min(a.cpu_sample, a.reads)
    if a.paired:
        a.cpu_sample &= ~1


# ----------------------------------------------------------------------------- workload
def workload_key(a):
    return "r%d_l%d_k%d_t%d_e%g%s_%s" % (a.reads, a.len, a.k, a.tx_bases, a.err, "_p" if a.paired else "", a.scaling)


def base_dir(a):
    return a.workdir or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")


def table_dir(a):
    return os.path.join(base_dir(a), "rcx_bench_" + hashlib.md5(workload_key(a).encode()).hexdigest()[:10])


def shard_dir(a, rseed, nreads):
    return os.path.join(base_dir(a), "rcx_bench_" + hashlib.md5(("%s_rs%d_n%d" % (workload_key(a), rseed, nreads)).encode()).hexdigest()[:10])


def synth_cmd(a, d, nreads, rseed, reads=True, table=False, dump=False, fq_limit=0):
    """One transcriptome (seed 1) for everything; the table's expected counts follow the WHOLE job's coverage (-cover),
    a shard draws `nreads` records with its own read seed."""
    ntx = max(4, a.tx_bases // 1400)
    per = 2 if a.paired else 1
    cmd = [SYNTH, "-seed", "1", "-k", str(a.k), "-ntx", str(ntx), "-txmin", "300", "-txmax", "2500", "-chimera", "0",
           "-polya", "0.3", "-sigma", "2.6", "-reads", str(nreads // per), "-cover", str(a.reads // per), "-len", str(a.len),
           "-err", str(a.err), "-nrate", "0.002", "-short", "0.02", "-rseed", str(rseed)]
    if a.paired:
        cmd += ["-paired"]
    if reads:
        cmd += ["-bin", os.path.join(d, "reads")]
    if fq_limit:
        cmd += ["-fq1", os.path.join(d, "sample.fq"), "-fqlimit", str(fq_limit // per)]
        if a.paired:
            cmd += ["-fq2", os.path.join(d, "sample2.fq")]
    if table:
        cmd += ["-table", os.path.join(d, "table")]
    if dump:
        cmd += ["-dump", os.path.join(d, "table.dump")]
    return cmd


def ensure_tools():
    # always through make (incremental): a tools/bin left by an older build would not know the generator's newer options
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")], stdout=sys.stderr)


def shard_plan(a, world, rank):
    """[(read seed, records)] of `rank`.  Weak scaling: every rank its own shard of a.reads records (read seed 1 + rank).
    Strong scaling: the job is ONE fixed set of NSUB sub-shards (read seeds 1..NSUB); rank r owns a contiguous run of them."""
    if a.scaling == "weak":
        return [(1 + rank, a.reads)]
    per = (a.reads // NSUB) & ~1
    sizes = [per] * (NSUB - 1) + [a.reads - per * (NSUB - 1)]
    lo, hi = NSUB * rank // world, NSUB * (rank + 1) // world
    return [(1 + s, sizes[s]) for s in range(lo, hi)]


def gen_shard(a, rseed, nreads):
    ensure_tools()
    d = shard_dir(a, rseed, nreads)
    os.makedirs(d, exist_ok=True)
    if not os.path.exists(os.path.join(d, "reads.ok")):
        t0 = time.time()
        subprocess.check_call(synth_cmd(a, d, nreads, rseed), stdout=sys.stderr)
        open(os.path.join(d, "reads.ok"), "w").close()
        log("[bench] generated %d reads (read seed %d) in %.1fs -> %s" % (nreads, rseed, time.time() - t0, d))
    return d


def gen_table(a, want_dump=False, fq_limit=0):
    """Table arrays (+ text dump and the FASTQ sample = the first reads of read seed 1, for the CPU legs)."""
    ensure_tools()
    d = table_dir(a)
    os.makedirs(d, exist_ok=True)
    need_table = not os.path.exists(os.path.join(d, "table.ok"))
    need_dump = want_dump and not os.path.exists(os.path.join(d, "table.dump.ok"))
    need_fq = fq_limit and not os.path.exists(os.path.join(d, "sample.fq.ok"))
    if need_table or need_dump or need_fq:
        t0 = time.time()
        nreads = max(fq_limit, 2)  # the reads are only drawn for the FASTQ sample
        subprocess.check_call(synth_cmd(a, d, nreads, 1, reads=False, table=True, dump=need_dump, fq_limit=fq_limit if need_fq else 0),
                              stdout=sys.stderr)
        open(os.path.join(d, "table.ok"), "w").close()
        if need_dump:
            open(os.path.join(d, "table.dump.ok"), "w").close()
        if need_fq:
            open(os.path.join(d, "sample.fq.ok"), "w").close()
        log("[bench] generated table%s%s in %.1fs" % (" + dump" if need_dump else "", " + sample" if need_fq else "", time.time() - t0))
    return d


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md 'clocks' line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def keep_dir(d):
    return os.path.join(d, "ref_keep")


def run_reference_binary(a, d, files, threads, poll=False, keep_first=False):
    """Run oracle/_ref/rcorrector once over `files` (-r each, or -p for a (mate1, mate2) tuple). Returns dict with the
    timed region ('Bad quality threshold' -> 'Processed') and, when poll=True, per-file completion times.  keep_first:
    the *.cor.fq of the first file (pair) and the scalars the binary printed are kept under ref_keep/ for the parity check."""
    outdir = os.path.join(d, "ref_out")
    os.makedirs(outdir, exist_ok=True)
    cor_name = lambda f: os.path.join(outdir, os.path.basename(f)[: os.path.basename(f).rfind(".")] + ".cor.fq")
    cmd = [REF_BIN]
    outs, extra_outs = [], []
    for f in files:
        if isinstance(f, tuple):
            cmd += ["-p", f[0], f[1]]
            outs.append(cor_name(f[1]))  # the mate-2 record of a pair is written last
            extra_outs.append(cor_name(f[0]))
        else:
            cmd += ["-r", f]
            outs.append(cor_name(f))
    cmd += ["-c", os.path.join(d, "table.dump"), "-k", str(a.k), "-od", outdir, "-t", str(threads), "-maxcorK", str(a.maxcork)]
    t_launch = time.time()
    p = subprocess.Popen(cmd, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL)
    marks = {}
    series = []
    stop = threading.Event()

    def poller():
        while not stop.is_set():
            now = time.time()
            sizes = []
            for o in outs:
                try:
                    sizes.append(os.path.getsize(o))
                except OSError:
                    sizes.append(0)
            series.append((now, sizes))
            time.sleep(0.002)

    th = None
    buf = b""
    text = ""
    while True:
        ch = p.stderr.read(1)  # stderr of the reference is unbuffered: timestamp as bytes arrive
        if not ch:
            break
        buf += ch
        if ch == b"\n":
            line = buf.decode(errors="replace")
            text += line
            buf = b""
            now = time.time()
            if line.startswith("Stored"):
                marks["stored"] = now
            elif line.startswith("Weak kmer"):
                marks["rate"] = now
            elif line.startswith("Bad quality"):
                marks["start"] = now
                if poll:
                    th = threading.Thread(target=poller, daemon=True)
                    th.start()
            elif line.startswith("Processed"):
                marks["end"] = now
                marks["reads"] = int(re.search(r"Processed (\d+) reads", line).group(1))
    p.wait()
    stop.set()
    if th:
        th.join()
    if p.returncode != 0 or "end" not in marks:
        raise RuntimeError("reference binary failed (rc %s)" % p.returncode)
    res = {"t_region": marks["end"] - marks["start"], "reads": marks["reads"], "t_load": marks["start"] - t_launch,
           "t_start": marks["start"], "t_end": marks["end"]}
    m = re.search(r"Weak kmer threshold rate: ([0-9.]+)", text)
    res["rate6"] = m.group(1) if m else None
    m = re.search(r"Bad quality threshold is '(.?)'", text, re.S)
    res["bad_qual"] = ord(m.group(1)) if m and m.group(1) else 0
    if poll:
        final = [os.path.getsize(o) for o in outs]
        done = []
        for i in range(len(outs)):
            t_done = marks["end"]
            for (t, sizes) in series:
                if sizes[i] >= final[i] - 4096 and final[i] > 0:
                    t_done = t
                    break
            done.append(t_done)
        done[-1] = marks["end"]
        res["done"] = done
    if keep_first:
        kd = keep_dir(d)
        os.makedirs(kd, exist_ok=True)
        first = [extra_outs[0], outs[0]] if isinstance(files[0], tuple) else [outs[0]]
        for i, f in enumerate(first):
            shutil.copyfile(f, os.path.join(kd, "sample%s.cor.fq" % ("" if i == 0 else "2")))
        json.dump({"rate6": res["rate6"], "bad_qual": res["bad_qual"], "reads": res["reads"] // len(files), "maxcork": a.maxcork, "k": a.k,
                   "threads": threads}, open(os.path.join(kd, "meta.json"), "w"))
    for o in outs + extra_outs:
        try:
            os.remove(o)
        except OSError:
            pass
    return res


def reference_sample_files(a, d, count):
    files = []
    for i in range(count):
        fs = []
        for src, tag in (("sample.fq", "s"), ("sample2.fq", "t"))[: 2 if a.paired else 1]:
            f = os.path.join(d, "sample_%s%02d.fq" % (tag, i))
            if not os.path.exists(f):
                os.link(os.path.join(d, src), f)
            fs.append(f)
        files.append(tuple(fs) if a.paired else fs[0])
    return files


def cpu_baseline(a, d):
    cores = os.cpu_count() or 1
    files = reference_sample_files(a, d, 1)
    r = run_reference_binary(a, d, files, cores, keep_first=True)
    return {"value": r["reads"] / r["t_region"], "unit": "reads/s", "cores": cores, "kind": "reference",
            "sample": "first %d reads of the workload, full %d-k-mer table, oracle/_ref/rcorrector -t %d, timed 'Bad quality threshold' -> "
                      "'Processed' (table load %.0fs excluded)" % (r["reads"], a.tx_bases, cores, r["t_load"])}


def workload_config(a):
    """Identical for both arms (the driver compares it)."""
    names = {"2": "configs[1]", "3": "configs[2]-shaped (48 M of the 200 M reads)", "5": "configs[4]-shaped (scaled)"}
    shape = names.get(a.config, "custom")
    how = ("every GPU its own shard of that many reads (weak scaling)" if a.scaling == "weak" else
           "one fixed read set of %d sub-shards split over the GPUs (strong scaling)" % NSUB)
    return {"workload": "%s: %d synthetic %s %d bp reads, k=%d, ~%d distinct k-mers; %s" % (
                shape, a.reads, "paired-end (interleaved)" if a.paired else "single-end", a.len, a.k, a.tx_bases, how),
            "reads": a.reads, "read_len": a.len, "k": a.k, "paired": bool(a.paired), "max_fix_per_k": a.maxcork,
            "table_kmers": a.tx_bases, "error_rate_injected": a.err,
            "l2_policy": "inputs (GBs of reads and table per step) far exceed the 126 MB L2"}


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/rcorrector not built"}))
        return 0
    d = gen_table(a, want_dump=True, fq_limit=a.cpu_sample)
    cores = os.cpu_count() or 1
    files = reference_sample_files(a, d, a.warmup + a.steps)
    r = run_reference_binary(a, d, files, cores, poll=True, keep_first=True)
    done = r["done"]
    t0 = done[a.warmup - 1] if a.warmup > 0 else r["t_start"]
    t1 = done[-1]
    per_file = r["reads"] // (a.warmup + a.steps)
    dt = max(t1 - t0, 1e-9)
    value = per_file * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC % a.k, "value": value, "unit": "reads/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(a),
        "step_note": "bounded sample: %d reads per step, full-size table" % per_file,
        "mbase_per_s": value * a.len / 1e6,
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference",
                         "sample": "%d reads per step x %d timed steps (after %d warm-up files) in ONE process of oracle/_ref/rcorrector -t %d; "
                                   "table load + rate estimation (%.0fs) outside the timed region" % (per_file, a.steps, a.warmup, cores, r["t_load"])},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


class StdoutToStderr:
    """stdout carries the one JSON line only: until restore(), file descriptor 1 points at stderr, so whatever libraries
    (NCCL prints a version banner on stdout when NCCL_DEBUG is set) or child processes write there lands on stderr."""

    def __init__(self):
        self.saved = None
        try:
            sys.stdout.flush()
            saved = os.dup(1)
            try:
                os.dup2(2, 1)
                self.saved = saved
            except OSError:
                os.close(saved)
        except OSError:
            pass  # no usable stderr: leave stdout alone

    def restore(self):
        if self.saved is None:
            return
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.saved = None


# ----------------------------------------------------------------------------- pre-pass scalars, as the binaries derive them
def dump_order_prefix(codes, counts):
    """The records in the order of the text dump rcx_synth writes (descending count, stable), as far as the ERROR_RATE
    pass can need them: it stops after 100 000 records whose best last-base variant reaches 1 000 (main.cpp:314-351), and
    every record with a count >= 1000 is one."""
    big = np.flatnonzero(counts >= 1000)
    if len(big) >= 100000:
        return codes[big[np.argsort(-counts[big].astype(np.int64), kind="stable")]]
    return codes[np.argsort(-counts.astype(np.int64), kind="stable")]


def bad_quality_threshold(qual, off, paired, nrec):
    """GetBadQuality (main.cpp:88-128) over the records the reference binary sees in the sample file(s): the first 1e6
    records of the first Reads object = mate 1 only for -p."""
    idx = np.arange(0, nrec, 2 if paired else 1)[:1000000]
    o = off.astype(np.int64)
    lens = o[idx + 1] - o[idx]
    first = np.bincount(qual[o[idx]].astype(np.int64), minlength=300)
    last = np.bincount(qual[o[idx] + np.maximum(lens, 1) - 1].astype(np.int64), minlength=300)
    total = len(idx)

    def cut(h):
        c = 0
        for i in range(300):
            c += int(h[i])
            if c > total * 0.05:
                return i
        return 300
    t1, t2 = cut(first) - 1, cut(last)
    v = min(t1, t2) & 0xff
    return v - 256 if v >= 128 else v


# ----------------------------------------------------------------------------- parity inside the bench
def parity_check(a, cor, td, mode, er, bq):
    """The sample reads through both host interfaces, formatted like Reads.h:360-421, against the reference's bytes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rcx_testlib as T  # checker only (FASTQ reader / writer restatement, oracle binding)
    from rcorrector_b200.api import apply_edits, quality_bits

    fq = [os.path.join(td, "sample.fq")] + ([os.path.join(td, "sample2.fq")] if a.paired else [])
    bs = [T.read_fastq(f) for f in fq]
    b = T.interleave(*bs) if a.paired else bs[0]
    kd = keep_dir(td)
    want, against = None, None
    meta_p = os.path.join(kd, "meta.json")
    if os.path.exists(meta_p):
        meta = json.load(open(meta_p))
        if meta.get("reads") == b.n and meta.get("maxcork") == a.maxcork and meta.get("k") == a.k:
            want = [open(os.path.join(kd, "sample%s.cor.fq" % s), "rb").read() for s in (("", "2") if a.paired else ("",))]
            against = "oracle/_ref/rcorrector (unmodified reference), *.cor.fq of this box's run"
            if meta.get("bad_qual") != bq or meta.get("rate6") != "%.6f" % er:
                raise SystemExit("bench: the arms ran different parameters: reference ERROR_RATE %s / bad quality %r, this arm %.6f / %r"
                                 % (meta.get("rate6"), meta.get("bad_qual"), er, bq))
    if want is None:  # no reference output on this box (N > 1 without the reference arm before it): the CPU oracle
        codes = np.fromfile(os.path.join(td, "table.codes"), dtype=np.uint64)
        counts = np.fromfile(os.path.join(td, "table.counts"), dtype=np.int32)
        table = T.OracleTable(a.k)
        table.put(codes, counts)
        so, corr, l, m, h = T.oracle_correct(table, b, a.k, mode, er, bq, a.maxcork, threads=os.cpu_count() or 1)
        want = [T.format_fastq(b, so, corr, l, m, h, rows=range(i, b.n, 2) if a.paired else None) for i in range(2 if a.paired else 1)]
        against = "oracle/ (C restatement pinned to the reference by tests/)"
    ifaces = []
    for name in ("rcx_correct_batch", "rcx_correct_batch_compact"):
        if name == "rcx_correct_batch":
            so, corr, l, m, h = cor.correct(mode, b.seq, b.qual, b.off)
        else:
            corr, l, m, h, edits = cor.correct_compact(mode, b.seq, quality_bits(b.qual, bq), b.off)
            so = apply_edits(b.seq, b.off, corr, edits)
        got = [T.format_fastq(b, so, corr, l, m, h, rows=range(i, b.n, 2) if a.paired else None) for i in range(2 if a.paired else 1)]
        if got != want:
            raise SystemExit("bench: PARITY FAILURE: %s output differs from %s on the %d-read sample" % (name, against, b.n))
        ifaces.append(name)
    return {"reads": b.n, "identical": True, "against": against, "interfaces": ifaces, "corrected_bases": int(corr[corr > 0].sum()),
            "bytes_compared": sum(len(w) for w in want)}


# ----------------------------------------------------------------------------- our arm
def measure(a, steps, warmup, world, rank, local, dist, guard_cpu_leg=True):
    """One workload on this rank's GPU; returns the JSON line (rank 0) or None."""
    import torch

    from rcorrector_b200 import INTERLEAVED, SINGLE, Corrector
    from rcorrector_b200.api import quality_bits
    from rcorrector_b200.sharding import broadcast_table

    mode = INTERLEAVED if a.paired else SINGLE
    dev = torch.device("cuda", local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    def sync_ms(t0):
        torch.cuda.synchronize()
        return 1e3 * (time.time() - t0)

    cor = Corrector(local)
    # ---- table: rank 0 generates / loads, NCCL broadcast, local insert on every GPU; pre-pass scalars like the binaries
    want_cpu = rank == 0 and world == 1 and not a.no_cpu_baseline and os.path.exists(REF_BIN) and guard_cpu_leg
    want_parity = not a.no_parity
    table_t = {}
    td = None
    if rank == 0:
        td = gen_table(a, want_dump=want_cpu, fq_limit=a.cpu_sample if (want_cpu or want_parity) else 0)
        codes = np.fromfile(os.path.join(td, "table.codes"), dtype=np.uint64)
        counts = np.fromfile(os.path.join(td, "table.counts"), dtype=np.int32)
    else:
        codes = counts = None
    if world > 1:
        t0 = time.time()
        d_codes, d_counts = broadcast_table(codes, counts, device=dev)
        table_t["broadcast_ms"] = sync_ms(t0)  # includes rank 0's host -> device upload of the 12 B / k-mer arrays
        nk = int(d_codes.numel())
        t0 = time.time()
        cor.build_table(a.k, d_codes, d_counts, n=nk, on_device=True)
        cor.sync()
        table_t["build_ms"] = sync_ms(t0)
        del d_codes, d_counts
    else:
        t0 = time.time()
        cor.build_table(a.k, codes, counts)
        cor.sync()
        table_t["build_ms"] = sync_ms(t0)  # host arrays: upload + insert
    info = cor.table_info()
    table_t.update({"kmers": info["entries"], "bytes": info["bytes"], "bytes_per_kmer": info["bytes"] / max(1, info["entries"])})

    # ---- this rank's reads
    plan = shard_plan(a, world, rank)
    dirs = [gen_shard(a, rs, nr) for rs, nr in plan]
    # sub-batches of at most 4 sub-shards (offsets are 32-bit: <= 4.29 G bases per batch)
    groups = [dirs[i:i + 4] for i in range(0, len(dirs), 4)]
    batches = []
    n_rank = 0
    windows = 0
    for g in groups:
        seqs = [np.fromfile(os.path.join(d, "reads.seq"), dtype=np.uint8) for d in g]
        quals = [np.fromfile(os.path.join(d, "reads.qual"), dtype=np.uint8) for d in g]
        offs = [np.fromfile(os.path.join(d, "reads.off"), dtype=np.uint32).astype(np.int64) for d in g]
        base = 0
        cat = [np.zeros(1, dtype=np.int64)]
        for o in offs:
            cat.append(o[1:] + base)
            base += int(o[-1])
        off = np.concatenate(cat)
        assert off[-1] < (1 << 32)
        seq_h = seqs[0] if len(seqs) == 1 else np.concatenate(seqs)
        qual_h = quals[0] if len(quals) == 1 else np.concatenate(quals)
        del seqs, quals
        lens = np.diff(off)
        windows += int(np.maximum(lens - a.k + 1, 0).sum())
        batches.append({"seq": seq_h, "qual": qual_h, "off": off.astype(np.uint32), "n": len(off) - 1, "total": int(off[-1]),
                        "max_len": int(lens.max())})
        n_rank += len(off) - 1

    # pre-pass scalars, derived like the binaries derive them (rank 0, broadcast): ERROR_RATE from the dump-order records
    # (main.cpp:310-358, GPU probes), badQualityThreshold from the sample's quality histogram (main.cpp:88-128)
    if rank == 0:
        er = cor.estimate_error_rate(dump_order_prefix(codes, counts), 0.95)
        b0 = batches[0]
        bq = bad_quality_threshold(b0["qual"], b0["off"], a.paired, min(a.cpu_sample, b0["n"]))
        scal = [er, float(bq)]
    else:
        scal = [0.0, 0.0]
    if world > 1:
        t = torch.tensor(scal, dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        scal = [float(x) for x in t.tolist()]
    er, bq = scal[0], int(scal[1])
    del codes, counts
    cor.set_params(a.k, er, bq, a.maxcork)

    pin = lambda x: torch.from_numpy(x).pin_memory()
    for b in batches:
        n, total = b["n"], b["total"]
        b["seq_p"], b["off_p"] = pin(b["seq"]), pin(b["off"])
        b["qbits_p"] = pin(quality_bits(b["qual"], bq))
        qual_p = pin(b["qual"])
        b["corr_p"] = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(4)]
        b["edits_p"] = torch.empty(max(total // 8, 1 << 20), dtype=torch.int16).pin_memory()
        # device-resident copy for the kernel-only arm
        b["d_seq0"] = b["seq_p"].to(dev, non_blocking=True)
        b["d_qual"] = qual_p.to(dev, non_blocking=True)
        b["d_off"] = b["off_p"].to(dev, non_blocking=True)
        b["d_seq"] = torch.empty_like(b["d_seq0"])
        b["d_res"] = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
        torch.cuda.synchronize()
        del qual_p
        b["seq"] = b["qual"] = None
    stream = torch.cuda.ExternalStream(cor.stream(), device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(collect=None):
        for b in batches:
            with torch.cuda.stream(stream):
                b["d_seq"].copy_(b["d_seq0"], non_blocking=True)  # restore the uncorrected reads (correction is in place)
            cor.correct_device(mode, b["n"], b["d_seq"], b["d_qual"], b["d_off"], b["total"], b["max_len"], *b["d_res"])
            if collect is not None:
                collect.append(cor.stats())

    edits_seen = [0]

    def step_e2e(collect=None):
        tot = 0
        for b in batches:
            out = cor.correct_compact(mode, b["seq_p"], b["qbits_p"], b["off_p"], out=(*b["corr_p"], b["edits_p"]))
            tot += len(out[4])
            if collect is not None:
                collect.append(cor.stats())
        edits_seen[0] = tot

    by_rank = []  # per timed region: ms per step of every rank (the line reports the maximum)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        cor.stats()  # (harvests the warm-up batches: the library picks the search kernel's mode from the previous batch's statistics)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        by_rank.append([ms / steps])
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            by_rank[-1] = [float(x.item()) / steps for x in g]
            ms = max(float(x.item()) for x in g)
        return ms

    rr32 = cor.measure_random_read(32)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(step_device, steps, warmup)
    ms_e2e = timed(step_e2e, steps, warmup)
    clocks = sampler.stop() if sampler else None
    # counters + per-kernel CUDA-event times: one more (untimed) step of each arm, summed over the sub-batches
    sd, se = [], []
    step_device(sd)
    step_e2e(se)
    cor.sync()
    acc = lambda lst: {k: sum(s[k] for s in lst) for k in lst[0]}
    st, st_e2e = acc(sd), acc(se)
    kern_by_rank = [[st["ms_scan"], st["ms_dfs"], st["ms_finalize"]]]
    if world > 1:
        t = torch.tensor(kern_by_rank[0], dtype=torch.float64, device=dev)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        kern_by_rank = [[round(float(v), 2) for v in x.tolist()] for x in g]

    # size-independent parity property inside the bench: both arms agree on every read's correction count
    chk = [sum(int(b["d_res"][0].sum().item()) for b in batches), sum(int(b["corr_p"][0].sum().item()) for b in batches)]
    if chk[0] != chk[1]:
        raise SystemExit("bench: device-resident and host arms disagree (%d vs %d corrected-base checksum)" % tuple(chk))
    if st["unsupported"]:
        raise SystemExit("bench: %d reads needed a case the CUDA path does not reproduce (rcx_stats.unsupported)" % st["unsupported"])

    cpu = None
    if want_cpu:
        try:
            cpu = cpu_baseline(a, td)  # (also leaves the reference's *.cor.fq of the sample for the parity check)
        except Exception as e:  # keep the GPU numbers even if the CPU leg fails
            cpu = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    parity = None
    if rank == 0 and want_parity:
        parity = parity_check(a, cor, td, mode, er, bq)
    if world > 1:
        dist.barrier()
    bytes_h2d = sum(b["total"] + (b["total"] + 7) // 8 + 4 * (b["n"] + 1) for b in batches)
    bytes_d2h = sum(16 * b["n"] for b in batches) + 2 * edits_seen[0]
    if world > 1:
        t = torch.tensor([n_rank, windows, bytes_h2d, bytes_d2h, st["kernel_launches"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        reads_total, _, h2d_all, d2h_all, launches_all = [int(x) for x in t.tolist()]
    else:
        reads_total, h2d_all, d2h_all, launches_all = n_rank, bytes_h2d, bytes_d2h, int(st["kernel_launches"])
    cor.close()
    if rank != 0:
        return None

    value = reads_total * steps / (ms_dev / 1e3)
    e2e_value = reads_total * steps / (ms_e2e / 1e3)
    # per-kernel CUDA-event times of one device step on rank 0; dominant kernel and its roofline.
    # Algorithmic bytes follow SURVEY §8d / DESIGN.md §5: one GetCount = one 32-byte sector.  The scan needs one GetCount per
    # k-window plus the bases; a search node needs the 4 one-base extensions of its k-mer; the l/m/h re-scan one per window
    # that covers a changed base.  "requests" is what the context-keyed table actually issues (one sector answers two
    # windows / a whole node); requests/s is compared with the random 32-byte request rate measured over this table.
    bases = st["bases"]
    kern = {
        "rcx_scan_thread_kernel": {"ms": st["ms_scan"], "requests": st["scan_probes"], "bytes": 32 * windows + bases,
                                   "does": "scan + GetStrongTrustedThreshold + triage + first-pass preparation"},
        "rcx_prep_kernel": {"ms": st["ms_prep"], "requests": 0, "bytes": 0, "does": "second-pass re-threshold + preparation (few reads)"},
        "rcx_dfs_kernel": {"ms": st["ms_dfs"], "requests": st["search_probes"], "nodes": st["search_nodes"],
                           "bytes": 32 * 4 * st["search_probes"] + 2 * bases, "does": "ordered correction search"},
        "rcx_finalize_thread_kernel": {"ms": st["ms_finalize"], "requests": st["rescan_probes"], "bytes": 32 * st["rescan_windows"] + 6 * bases,
                                       "does": "post-filters + apply + l/m/h"},
    }
    req_peak = rr32 / 32.0  # G requests/s
    for v in kern.values():
        v["gbs"] = v["bytes"] / max(v["ms"], 1e-9) / 1e6
        v["grequests_per_s"] = v["requests"] / max(v["ms"], 1e-9) / 1e6
        v["frac_of_random_request_peak"] = v["grequests_per_s"] / req_peak
        v["dram_line_model_gbs"] = 128 * v["requests"] / max(v["ms"], 1e-9) / 1e6  # an L2 miss drags a 128-byte line (profiles/r01_f)
    dom = max(kern, key=lambda k: kern[k]["ms"])
    dom_ms, dom_bytes = kern[dom]["ms"], kern[dom]["bytes"]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    scan = kern["rcx_scan_thread_kernel"]
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj["workload"] == {"reads": n_rank, "len": a.len, "k": a.k}:
            traffic = {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}
    except Exception:
        pass
    line = {
        "metric": METRIC % a.k, "value": value, "unit": "reads/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(a),
        "mbase_per_s": value * a.len / 1e6,
        "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": ms_e2e / steps, "interface": "rcx_correct_batch_compact",
                "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "link_gbs_per_gpu": {"h2d": bytes_h2d / (ms_e2e / steps) / 1e6, "d2h": bytes_d2h / (ms_e2e / steps) / 1e6},
                # rank 0, summed over the chunks of one step; chunks overlap, so the sum exceeds ms_per_step (a copy's time includes
                # its wait for the copy engine)
                "chunk_sums_ms": {k: round(float(st_e2e["ms_" + k]), 2) for k in ("h2d", "scan", "dfs", "finalize", "d2h", "total")}},
        "gpu_launches": int(launches_all) * steps,  # kernels of the timed region of `value`, all GPUs (the e2e region adds 3 per chunk for the edits)
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic.get(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": int(dom_bytes / max(1, len(batches))),
                     "launches_per_step": len(batches), "kernel_ms": dom_ms,
                     "random_request_peak": {"grequests_per_s": req_peak, "gbs_at_32B": rr32,
                                             "how": "independent random 32-byte reads over this table, measured in this run"},
                     "frac_of_random_request_peak": kern[dom]["frac_of_random_request_peak"],
                     "lookup_kernel": {"kernel": "rcx_scan_thread_kernel", "achieved": scan["gbs"], "frac": scan["gbs"] / hbm_peak,
                                       "grequests_per_s": scan["grequests_per_s"],
                                       "frac_of_random_request_peak": scan["frac_of_random_request_peak"], "kernel_ms": scan["ms"],
                                       "algorithmic_bytes_per_step": int(scan["bytes"]), "requests": int(scan["requests"]),
                                       "traffic": traffic.get("rcx_scan_thread_kernel")},
                     "note": "random-access bound: 32 B per GetCount (one DRAM sector) + streamed bases against the streaming-copy peak; "
                             "the request fractions compare sector requests/s with the measured random-request rate (each miss moves a 128-B line)"},
        "kernels": kern, "work_items": st["work_items"], "second_pass": st["second_pass"],
        "table": table_t, "params": {"error_rate": er, "bad_qual": bq, "source": "rcx_estimate_error_rate + quality histogram (as the binaries)"},
        "reads_this_rank": n_rank, "clocks": clocks,
        "by_rank": {"ms_per_step": [round(v, 2) for v in by_rank[0]], "e2e_ms_per_step": [round(v, 2) for v in by_rank[1]],
                    "scan_dfs_finalize_ms": kern_by_rank},
    }
    if parity is not None:
        line["parity"] = parity
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def main_ours(a):
    guard = StdoutToStderr()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    line = measure(a, a.steps, a.warmup, world, rank, local, dist)
    if world == 1 and a.config == "2" and not a.no_scaling_base:
        # the base of the strong-scaling curve the N > 1 runs draw: the configs[2]-shaped job on this one GPU.  In a child
        # process (this one's device memory is released first), so that nothing it does can cost the headline line.
        try:
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            cmd = [sys.executable, os.path.abspath(__file__), "--config", "3", "--no-cpu-baseline", "--no-parity", "--no-scaling-base",
                   "--steps", str(min(a.steps, 5)), "--warmup", str(min(a.warmup, 3))] + (["--workdir", a.workdir] if a.workdir else [])
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local)))
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=sys.stderr, timeout=600, env=env)
            l3 = json.loads(r.stdout.decode().strip().splitlines()[-1])
            line["scaling_base_cfg3"] = {k: l3[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "mbase_per_s", "table")}
            line["scaling_base_cfg3"]["e2e"] = {k: l3["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
            line["scaling_base_cfg3"]["kernels_ms"] = {k: v["ms"] for k, v in l3["kernels"].items()}
            line["scaling_base_cfg3"]["lookup_kernel"] = l3["roofline"]["lookup_kernel"]
        except BaseException as e:  # the headline line stands on its own
            line["scaling_base_cfg3"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    guard.restore()
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    args = parse_args()
    sys.exit(main_reference(args) if args.impl == "reference" else main_ours(args))
