#!/bin/bash
# Host-side throughput of the drop-in binary WITHOUT a GPU: rcorrector_main.cpp linked against the test double of librcx
# (tests/harness/rcx_double.cpp) in pass-through mode (no correction at all), so that what is timed is the binary's own
# read / parse / format / write pipeline.  Test infrastructure; nothing here is shipped.
# usage: bash tools/host_rate.sh [reads=4000000] [threads="1 4 8"]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${RCX_HOST_RATE_DIR:-/dev/shm/rcx_host_rate}
N=${1:-4000000}
THREADS=${2:-"1 4 8"}
mkdir -p $W
make -s -C $ROOT/tools && make -s -C $ROOT/oracle librcx_oracle.so
[ -f $W/r_$N.fq ] || $ROOT/tools/bin/rcx_synth -seed 1 -k 23 -ntx 2000 -reads $N -len 100 -err 0.01 -fq1 $W/r_$N.fq -dump $W/t.dump 2>/dev/null
g++ -O2 -std=c++17 -pthread -o $W/rcorrector $ROOT/rcorrector_b200/csrc/rcorrector_main.cpp $ROOT/tests/harness/rcx_double.cpp \
    -L$ROOT/oracle -l:librcx_oracle.so -lz -Wl,-rpath,$ROOT/oracle
for t in $THREADS; do
  rm -rf $W/out
  echo "== -t $t"
  RCX_DOUBLE_PASSTHROUGH=1 RCX_TIMING=1 $W/rcorrector -r $W/r_$N.fq -c $W/t.dump -k 23 -od $W/out -t $t 2>&1 | grep "pinned batch\|done\|stage sums"
done
