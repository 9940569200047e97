#!/bin/bash
# last measurements: ncu --set full of the three kernels on the configs[2] shape (150-bp pairs), and the k-mer counting rate
TAG=${1:-r02_r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"rcx_(scan_thread|dfs|finalize_thread)" -s 8 -c 4 -f -o $OUT/full_cfg3 \
  python bench.py --config 3 --no-cpu-baseline --no-parity --no-scaling-base --steps 1 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
python - <<'P' > $OUT/count_rate.json 2> $OUT/count_rate.err
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from rcorrector_b200 import Corrector
a = bench.parse_args(["--config", "2"])
(rs, nr), = bench.shard_plan(a, 1, 0)
d = bench.gen_shard(a, rs, nr)
seq = np.fromfile(os.path.join(d, "reads.seq"), dtype=np.uint8)
off = np.fromfile(os.path.join(d, "reads.off"), dtype=np.uint32)
cor = Corrector(0)
res = {}
for hint in (0, 400_000_000):
    t0 = time.time()
    codes, counts, distinct, inst = cor.count_kmers(a.k, [(seq, off)], capacity_hint=hint)
    dt = time.time() - t0
    res["capacity_hint_%d" % hint] = {"seconds": round(dt, 3), "reads": len(off) - 1, "kmer_instances": inst, "distinct": distinct,
                                       "kept_count_ge_2": len(codes), "instances_per_s": inst / dt, "table": cor.table_info()}
print(json.dumps(res))
P
cat $OUT/count_rate.json | cut -c1-600; tail -2 $OUT/count_rate.err
