#!/bin/bash
# Compare compile-time variants of one kernel file in ONE gpurun call (bench.py picks the library up through RCX_LIB).
#   build here (no GPU needed):   bash tools/variant_round.sh build rcx_dfs.cu mb7:-DDFS_MINBLOCKS=7 ml12:-DDFS_MINLANES=12 ...
#   run on the box:               gpurun -- 'bash tools/variant_round.sh run <tag> mb7 ml12 ...'
# Results so far (profiles/r01_u_e2e_chunks.md): DFS_MINBLOCKS 7 / 8 equal / slower than 6 (r01_y).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/rcorrector_b200/csrc
mode=$1; shift
if [ "$mode" == "build" ]; then
  src=$1; shift
  base=${src%.cu}
  mkdir -p $ROOT/build
  make -s -C $CS lib
  others=$(cd $CS && ls rcx_*.o | grep -v "^$base.o$" | sed "s#^#$CS/#")
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c $CS/$src -o $ROOT/build/${base}_$name.o
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/librcx_$name.so $others $ROOT/build/${base}_$name.o -lcudart
    echo "built build/librcx_$name.so ($flags)"
  done
  exit 0
fi
TAG=$1; shift
OUT=$ROOT/gpurun_out/$TAG; mkdir -p $OUT
A="--no-cpu-baseline --steps 3 --warmup 3"
P="--paired --reads 6000000 --len 150 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
Q="--reads 4000000 --len 150 --k 31 --err 0.04 --maxcork 8 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
cd $ROOT
for v in default "$@"; do
  lib=$ROOT/rcorrector_b200/librcx.so; [ "$v" != default ] && lib=$ROOT/build/librcx_$v.so
  RCX_LIB=$lib timeout 200 python bench.py $A > $OUT/bench_100_$v.json 2> $OUT/bench_100_$v.err
  RCX_LIB=$lib timeout 200 python bench.py $P > $OUT/bench_150p_$v.json 2> $OUT/bench_150p_$v.err
  RCX_LIB=$lib timeout 200 python bench.py $Q > $OUT/bench_k31_$v.json 2> $OUT/bench_k31_$v.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.1f ms/step"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms"],1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
