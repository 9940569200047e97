#!/bin/bash
# One gpurun call: bench lines of the other BASELINE shapes (scaled to one GPU and a short call), K1 long-read variants.
# usage (repo root, GPU box): bash tools/shapes_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
P="--paired --reads 6000000 --len 150 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
Q="--reads 4000000 --len 150 --k 31 --err 0.04 --maxcork 8 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
for v in 0 1; do
  RCX_TS_LONG=$v timeout 240 python bench.py $P > $OUT/bench_150p_v$v.json 2> $OUT/bench_150p_v$v.err; echo "150p v$v rc=$?"
  RCX_TS_LONG=$v timeout 240 python bench.py $Q > $OUT/bench_150_k31_v$v.json 2> $OUT/bench_150_k31_v$v.err; echo "k31 v$v rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "%.1f ms/step"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms"],1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
