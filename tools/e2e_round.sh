#!/bin/bash
# One gpurun call: the host-buffer (e2e) arm against slots / chunk size / chunk chaining, two read shapes.
# usage (repo root, GPU box): bash tools/e2e_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
A="--no-cpu-baseline --steps 4 --warmup 3"
B="--paired --reads 6000000 --len 150 --tx-bases 50000000 --no-cpu-baseline --steps 4 --warmup 3"
for cfg in "3 0 200e6" "4 0 200e6" "3 1 200e6" "4 1 200e6" "3 2 200e6" "4 1 150e6" "4 1 100e6" "4 0 150e6"; do
  set -- $cfg
  RCX_SLOTS=$1 RCX_CHAIN=$2 RCX_CHUNK_BASES=$3 timeout 240 python bench.py $A > $OUT/e2e_100_s$1_c$2_b$3.json 2> $OUT/e2e_100_s$1_c$2_b$3.err
  RCX_SLOTS=$1 RCX_CHAIN=$2 RCX_CHUNK_BASES=$3 timeout 240 python bench.py $B > $OUT/e2e_150p_s$1_c$2_b$3.json 2> $OUT/e2e_150p_s$1_c$2_b$3.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/e2e_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "dev %.1f ms"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["ms_per_step"], d["e2e"].get("chunk_sums_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
