#!/bin/bash
# One gpurun call: GPU tests (optional) + bench lines of the default workload and of the other BASELINE shapes (scaled to
# one GPU and a short call).   usage (repo root, GPU box): bash tools/shapes_round.sh <tag> [tests]
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
fi
A="--no-cpu-baseline --steps 4 --warmup 3"
P="--paired --reads 6000000 --len 150 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
Q="--reads 4000000 --len 150 --k 31 --err 0.04 --maxcork 8 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
timeout 240 python bench.py $A > $OUT/bench_100.json 2> $OUT/bench_100.err; echo "100 rc=$?"
timeout 240 python bench.py $P > $OUT/bench_150p.json 2> $OUT/bench_150p.err; echo "150p rc=$?"
timeout 240 python bench.py $Q > $OUT/bench_150_k31.json 2> $OUT/bench_150_k31.err; echo "k31 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.1f ms/step"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms"],1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
