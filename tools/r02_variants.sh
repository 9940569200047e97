#!/bin/bash
# Compare compile-time variants of one kernel file in ONE gpurun call (RCX_LIB picks the library).
#   build here:  bash tools/r02_variants.sh build rcx_dfs.cu a5:"-DDFS_MINBLOCKS=5" a0:"-DDFS_AHEAD=0"
#   on the box:  bash tools/r02_variants.sh run <tag> "<configs>" a5 a0      (configs e.g. "2 3 5")
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/rcorrector_b200/csrc
mode=$1; shift
if [ "$mode" == "build" ]; then
  src=$1; shift
  base=${src%.cu}
  mkdir -p $ROOT/build
  make -s -C $CS lib || exit 1
  others=$(cd $CS && ls rcx_*.o | grep -v "^$base.o$" | sed "s#^#$CS/#")
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c $CS/$src -o $ROOT/build/${base}_$name.o || exit 1
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/librcx_$name.so $others $ROOT/build/${base}_$name.o -lcudart || exit 1
    echo "built build/librcx_$name.so ($flags)"
  done
  exit 0
fi
TAG=$1; shift
CFGS=$1; shift
OUT=$ROOT/gpurun_out/$TAG; mkdir -p $OUT
cd $ROOT
for cfg in $CFGS; do
  for v in default "$@"; do
    lib=$ROOT/rcorrector_b200/librcx.so; [ "$v" != default ] && lib=$ROOT/build/librcx_$v.so
    RCX_LIB=$lib timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-parity --no-scaling-base --steps 5 --warmup 3 > $OUT/bench_c${cfg}_$v.json 2> $OUT/bench_c${cfg}_$v.err
    python - $OUT/bench_c${cfg}_$v.json "cfg $cfg $v" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ms/step %.2f e2e %.2f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()), "| dfs req %d" % d["kernels"]["rcx_dfs_kernel"]["requests"])
except Exception as e:
    print(sys.argv[2], "unreadable", e)
P
  done
done
